// contract.cuh -- device replacement of btas::contract (see contract.cu).
#pragma once
#include "common.cuh"

namespace qcp2 {

void contract(Ctx &ctx, double alpha, const TView &A, const char *ia, const TView &B, const char *ib, double beta,
              const TView &C, const char *ic, const double *cin = nullptr);
// cin (optional, same layout as C): C = beta*cin + alpha*A.B, i.e. the "+= integrals" terms of the
// reference (cc.cpp:232, 240, 250, 312) folded into the GEMM epilogue instead of a separate sweep

}// namespace qcp2
