// contract.cuh -- device replacement of btas::contract (see contract.cu).
#pragma once
#include "common.cuh"

namespace qcp2 {

void contract(Ctx &ctx, double alpha, const TView &A, const char *ia, const TView &B, const char *ib, double beta,
              const TView &C, const char *ic);

}// namespace qcp2
