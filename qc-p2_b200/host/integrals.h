// integrals.h -- same declarations as the reference's src/integrals.h:11-34; the bodies in
// integrals.cpp forward to the C ABI (include/qcp2_b200.h) instead of btas::contract / loops.
#pragma once
#include <string>

#include "basis.h"
#include "gaussian.h"
#include "general.h"

struct int_struct {// integrals.h:11-23
    DTensor oooo, ovoo, oovo, ooov, ovov, ovvo, oovv, vovv, ovvv, vvvo, vvvv;
};

// DTensor eri_ao_tensor(const BasisSet &obs);  -> gaussian.h (host, boundary input of the hot path)
DTensor transform_ao_mo(const DTensor &pq_rs, const Matrix &Coeff1, const Matrix &Coeff2);
DTensor make_denom(const DTensor &F_spin, int no, int nv);
DTensor transform_to_so(const DTensor &mo_ints_aa, const DTensor &mo_ints_bb, const DTensor &mo_ints_ab);
DTensor slice_ints(const DTensor &so_ints, int no, int nv, std::string int_string);
