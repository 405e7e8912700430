// backend.cpp -- TEST INFRASTRUCTURE ONLY (oracle/ref_shim); see backend.h.
#include "backend.h"

#include <cstring>
#include <iostream>

#include "../../qc-p2_b200/host/basis.h"
#include "../../qc-p2_b200/host/gaussian.h"

#define SHIMB_API extern "C" __attribute__((visibility("default")))

static std::vector<Atom> atoms_of(int natom, const int *Z, const double *xyz) {
    std::vector<Atom> a;
    for (int k = 0; k < natom; ++k) a.push_back(Atom{Z[k], xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2]});
    return a;
}

SHIMB_API void *shimb_basis_create(const char *name, int natom, const int *Z, const double *xyz) {
    try {
        return new BasisSet(name, atoms_of(natom, Z, xyz));
    } catch (const std::exception &e) {
        std::cerr << "ref_shim backend: " << e.what() << std::endl;
        return nullptr;
    }
}
SHIMB_API void shimb_basis_destroy(void *h) { delete static_cast<BasisSet *>(h); }
SHIMB_API int shimb_nshell(void *h) { return (int) static_cast<BasisSet *>(h)->shells().size(); }
SHIMB_API int shimb_nbf(void *h) { return (int) static_cast<BasisSet *>(h)->nbf(); }
SHIMB_API int shimb_shell(void *h, int s, int *l, double *origin, double *alpha, double *coef, int maxprim) {
    const Shell &sh = static_cast<BasisSet *>(h)->shells()[(size_t) s];
    *l = sh.l;
    for (int d = 0; d < 3; ++d) origin[d] = sh.O[(size_t) d];
    const int np = (int) sh.alpha.size();
    for (int k = 0; k < np && k < maxprim; ++k) alpha[k] = sh.alpha[(size_t) k], coef[k] = sh.coef[(size_t) k];
    return np;
}
SHIMB_API int shimb_onebody(void *h, int op, int natom, const int *Z, const double *xyz, double *out) {
    try {
        const BasisSet &b = *static_cast<BasisSet *>(h);
        const Operator1 o = op == 0 ? Operator1::overlap : op == 1 ? Operator1::kinetic : Operator1::nuclear;
        Matrix m = compute_1body_ints(b, o, atoms_of(natom, Z, xyz));
        std::memcpy(out, m.data(), sizeof(double) * b.nbf() * b.nbf());
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "ref_shim backend: " << e.what() << std::endl;
        return 1;
    }
}
SHIMB_API int shimb_eri(void *h, double *out) {
    try {
        DTensor t = eri_ao_tensor(*static_cast<BasisSet *>(h));
        std::memcpy(out, t.data(), sizeof(double) * t.size());
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "ref_shim backend: " << e.what() << std::endl;
        return 1;
    }
}
