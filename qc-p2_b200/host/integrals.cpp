// integrals.cpp -- integrals.h entry points on the GPU (reference bodies: integrals.cpp:95-207).
#include "integrals.h"

#include "gpu.h"

DTensor transform_ao_mo(const DTensor &pq_rs, const Matrix &Coeff1, const Matrix &Coeff2) {
    const int n = (int) pq_rs.extent(0);
    DTensor ij_kl(n, n, n, n);
    gpu_check(qcp2_transform_ao_mo(gpu(), pq_rs.data(), Coeff1.data(), Coeff2.data(), n, ij_kl.data()));
    return ij_kl;
}

DTensor make_denom(const DTensor &F_spin, int no, int nv) {
    DTensor denom(no, no, nv, nv);
    gpu_check(qcp2_make_denom(gpu(), F_spin.data(), (int) F_spin.extent(0), no, nv, denom.data()));
    return denom;
}

DTensor transform_to_so(const DTensor &aa, const DTensor &bb, const DTensor &ab) {
    const int n = (int) aa.extent(0);
    DTensor so(2 * n, 2 * n, 2 * n, 2 * n);
    gpu_check(qcp2_transform_to_so(gpu(), aa.data(), bb.data(), ab.data(), n, so.data()));
    return so;
}

DTensor slice_ints(const DTensor &so_ints, int no, int nv, std::string s) {
    if (s.size() != 4) throw std::invalid_argument("slice_ints: int_string must have 4 characters");
    auto e = [&](char c) { return c == 'o' ? no : nv; };
    DTensor out(e(s[0]), e(s[1]), e(s[2]), e(s[3]));
    gpu_check(qcp2_slice_ints(gpu(), so_ints.data(), (int) so_ints.extent(0), no, nv, s.c_str(), out.data()));
    return out;
}
