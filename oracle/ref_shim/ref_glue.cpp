// ref_glue.cpp -- TEST INFRASTRUCTURE ONLY (oracle/ref_shim): flat C entry points around the
// reference's OWN functions.  This file is compiled together with the UNMODIFIED reference
// sources /root/reference/src/{integrals,mp2,cc,hf,general}.cpp (against the API shims in
// include/) into oracle/_ref/libqcp2_ref.so.  It contains marshalling only -- raw pointers into
// DTensor / Matrix / Vector and back -- plus ref_post_hf, which strings the reference's functions
// together the way MP2() (mp2.cpp:45-86) and main() (main.cpp:76-88) do, minus the Libint call,
// so that seeded pseudo-molecules (random AO tensors) can be pushed through the reference too.
// Signatures mirror oracle.cpp's orc_* so the Python front end is shared.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unistd.h>

#include "src/cc.h"
#include "src/general.h"
#include "src/hf.h"
#include "src/integrals.h"
#include "src/mp2.h"

// not declared in cc.h but defined in cc.cpp:8-23
int_struct get_integrals(const scf_results &SCF, const mp2_results &MP2);

namespace {

DTensor t2(long a, long b, const double *p) {
    DTensor t(a, b);
    if (p) std::memcpy(t.data(), p, sizeof(double) * t.size());
    return t;
}
DTensor t4(long a, long b, long c, long d, const double *p) {
    DTensor t(a, b, c, d);
    if (p) std::memcpy(t.data(), p, sizeof(double) * t.size());
    return t;
}
void put(const DTensor &t, double *dst) {
    if (dst) std::memcpy(dst, t.data(), sizeof(double) * t.size());
}
Matrix mat(long r, long c, const double *p) {
    Matrix m(r, c);
    std::memcpy(m.data(), p, sizeof(double) * (size_t) (r * c));
    return m;
}
Vector vec(long n, const double *p) {
    Vector v(n);
    for (long k = 0; k < n; ++k) v(k) = p[k];
    return v;
}
const char *kIntNames[11] = {"oooo", "ovoo", "oovo", "ooov", "ovov", "ovvo", "oovv", "vovv", "ovvv", "vvvo", "vvvv"};
DTensor *int_field(int_struct &s, int k) {
    DTensor *f[11] = {&s.oooo, &s.ovoo, &s.oovo, &s.ooov, &s.ovov, &s.ovvo, &s.oovv, &s.vovv, &s.ovvv, &s.vvvo, &s.vvvv};
    return f[k];
}
int_struct wrap_ints(const double *const *p, long o, long v) {
    int_struct s;
    for (int k = 0; k < 11; ++k) {
        long e[4];
        for (int d = 0; d < 4; ++d) e[d] = kIntNames[k][d] == 'o' ? o : v;
        if (p[k]) *int_field(s, k) = t4(e[0], e[1], e[2], e[3], p[k]);
    }
    return s;
}
moes wrap_moes(const double *foo, const double *fov, const double *fvv, long o, long v) {
    moes m;
    m.F_ii = t2(o, o, foo);
    m.F_ia = t2(o, v, fov);// fov == NULL -> zeros
    m.F_aa = t2(v, v, fvv);
    return m;
}
cc_intermediates wrap_inter(const double *const *p, long o, long v) {
    cc_intermediates X;
    X.F_ae = t2(v, v, p[0]), X.F_mi = t2(o, o, p[1]), X.F_me = t2(o, v, p[2]);
    X.W_mnij = t4(o, o, o, o, p[3]), X.W_mbej = t4(o, v, v, o, p[4]), X.W_abef = t4(v, v, v, v, p[5]);
    return X;
}

// swallow the reference's std::cout / printf chatter (and count the CCSD iteration lines)
struct Quiet {
    int saved = -1;
    std::string path;
    std::streambuf *old = nullptr;
    std::ostringstream sink;
    Quiet() {
        std::fflush(stdout);
        char tmpl[] = "/tmp/qcp2_ref_XXXXXX";
        const int fd = mkstemp(tmpl);
        path = tmpl;
        saved = dup(1);
        dup2(fd, 1);
        close(fd);
        old = std::cout.rdbuf(sink.rdbuf());
    }
    std::string finish() {
        if (saved < 0) return "";
        std::fflush(stdout);
        dup2(saved, 1);
        close(saved);
        saved = -1;
        std::cout.rdbuf(old);
        std::ifstream f(path);
        std::stringstream ss;
        ss << f.rdbuf();
        unlink(path.c_str());
        return ss.str() + sink.str();
    }
    ~Quiet() { finish(); }
};
int count_cc_iterations(const std::string &log) {// lines " %02d %20.12f %20.12f" (cc.cpp:437)
    std::istringstream is(log);
    std::string line;
    int last = -1;
    while (std::getline(is, line)) {
        int it;
        double e, de;
        char extra;
        if (std::sscanf(line.c_str(), " %d %lf %lf %c", &it, &e, &de, &extra) == 3) last = it;
    }
    return last;
}
}// namespace

#define REF_TRY try {
#define REF_CATCH                                                   \
    }                                                               \
    catch (const std::exception &e) {                               \
        std::fprintf(stderr, "reference (oracle/_ref) error: %s\n", e.what()); \
        return 1;                                                   \
    }                                                               \
    return 0;

extern "C" {

// the shim's btas::contract itself (label strings, one character per index), so that the restated third-party
// semantics can be held against an independent einsum (tests/test_ref_pins_oracle.py)
int ref_shim_contract(double alpha, const double *A, const long *extA, const char *ia, const double *B, const long *extB,
                      const char *ib, double beta, double *C, const long *extC, const char *ic) {
    REF_TRY
    auto make = [](const double *p, const long *e, const char *labels) {
        const size_t r = std::strlen(labels);
        DTensor t;
        switch (r) {
            case 2: t = DTensor(e[0], e[1]); break;
            case 4: t = DTensor(e[0], e[1], e[2], e[3]); break;
            case 6: t = DTensor(e[0], e[1], e[2], e[3], e[4], e[5]); break;
            default: throw std::invalid_argument("ref_shim_contract: rank must be 2, 4 or 6");
        }
        std::memcpy(t.data(), p, sizeof(double) * t.size());
        return t;
    };
    auto labels = [](const char *s) {
        std::vector<int> v;
        for (; *s; ++s) v.push_back((int) *s);
        return v;
    };
    DTensor a = make(A, extA, ia), b = make(B, extB, ib), c = make(C, extC, ic);
    btas::contract_labels(alpha, a, labels(ia), b, labels(ib), beta, c, labels(ic));
    put(c, C);
    REF_CATCH
}

int ref_transform_ao_mo(const double *ao, const double *C1, const double *C2, int n, double *out) {
    REF_TRY
    put(transform_ao_mo(t4(n, n, n, n, ao), mat(n, n, C1), mat(n, n, C2)), out);
    REF_CATCH
}
int ref_transform_to_so(const double *aa, const double *bb, const double *ab, int n, double *out) {
    REF_TRY
    put(transform_to_so(t4(n, n, n, n, aa), t4(n, n, n, n, bb), t4(n, n, n, n, ab)), out);
    REF_CATCH
}
int ref_slice_ints(const double *so, int n2, int no, int nv, const char *spec, double *out) {
    REF_TRY
    put(slice_ints(t4(n2, n2, n2, n2, so), no, nv, spec), out);
    REF_CATCH
}
int ref_make_so_moes(const double *ea, const double *eb, int nao, double *out) {
    REF_TRY
    put(make_so_moes(vec(nao, ea), vec(nao, eb), nao), out);
    REF_CATCH
}
int ref_make_denom(const double *F_spin, int n2, int no, int nv, double *out) {
    REF_TRY
    put(make_denom(t2(n2, n2, F_spin), no, nv), out);
    REF_CATCH
}
int ref_run_mp2(const double *oovv, const double *denom, int no, int nv, double *T, double *E) {
    REF_TRY
    mp2_output r = run_mp2(t4(no, no, nv, nv, oovv), t4(no, no, nv, nv, denom));
    *E = r.energy;
    put(r.T, T);
    REF_CATCH
}
// e1/e2 have nvec entries (make_fock reads eps(n1/2 + i/2), cc.cpp:25-34)
int ref_make_fock(const double *e1, const double *e2, int nvec, int n1, int n2, double *out) {
    REF_TRY
    put(make_fock(vec(nvec, e1), vec(nvec, e2), n1, n2), out);
    REF_CATCH
}
// kind 0: D_ia, 1: D_ijab, 2: D_triples (cc.cpp:73-123)
int ref_make_D(const double *foo, const double *fvv, int no, int nv, int kind, double *out) {
    REF_TRY
    moes m = wrap_moes(foo, nullptr, fvv, no, nv);
    put(kind == 0 ? make_D_ia(m) : kind == 1 ? make_D_ijab(m) : make_D_triples(m), out);
    REF_CATCH
}
int ref_make_tau(const double *Ts, const double *Td, int no, int nv, int bar, double *out) {
    REF_TRY
    DTensor ts = t2(no, nv, Ts), td = t4(no, no, nv, nv, Td);
    put(bar ? make_tau_bar(ts, td) : make_tau(ts, td), out);
    REF_CATCH
}
int ref_multiply_Ts(const double *Ts, int no, int nv, double *out) {
    REF_TRY
    put(multiply_Ts(t2(no, nv, Ts)), out);
    REF_CATCH
}
int ref_update_intermediates(const double *Ts, const double *Td, const double *const *ints, const double *foo,
                             const double *fov, const double *fvv, int no, int nv, double *const *out) {
    REF_TRY
    cc_intermediates X = update_intermediates(t2(no, nv, Ts), t4(no, no, nv, nv, Td), wrap_ints(ints, no, nv),
                                              wrap_moes(foo, fov, fvv, no, nv));
    put(X.F_ae, out[0]), put(X.F_mi, out[1]), put(X.F_me, out[2]);
    put(X.W_mnij, out[3]), put(X.W_mbej, out[4]), put(X.W_abef, out[5]);
    REF_CATCH
}
int ref_make_T1(const double *Ts, const double *Td, const double *const *ints, const double *const *inter,
                const double *foo, const double *fov, const double *fvv, int no, int nv, double *out) {
    REF_TRY
    moes mo = wrap_moes(foo, fov, fvv, no, nv);
    put(make_T1(t2(no, nv, Ts), t4(no, no, nv, nv, Td), wrap_ints(ints, no, nv), wrap_inter(inter, no, nv), make_D_ia(mo), mo), out);
    REF_CATCH
}
int ref_make_T2(const double *Ts, const double *Td, const double *const *ints, const double *const *inter,
                const double *foo, const double *fvv, int no, int nv, double *out) {
    REF_TRY
    moes mo = wrap_moes(foo, nullptr, fvv, no, nv);
    put(make_T2(t2(no, nv, Ts), t4(no, no, nv, nv, Td), wrap_ints(ints, no, nv), wrap_inter(inter, no, nv), make_D_ijab(mo)), out);
    REF_CATCH
}
int ref_ccsd_energy(const double *Ts, const double *Td, const double *oovv, const double *fov, int no, int nv, double *E) {
    REF_TRY
    moes mo;
    mo.F_ia = t2(no, nv, fov);
    *E = ccsd_energy(t2(no, nv, Ts), t4(no, no, nv, nv, Td), t4(no, no, nv, nv, oovv), mo);
    REF_CATCH
}
// CCSD_T (cc.cpp:454-560) on the amplitudes / blocks given; the other blocks stay empty
int ref_ccsd_t(const double *T1, const double *T2, const double *oovv, const double *vovv, const double *ovoo,
               const double *foo, const double *fvv, int no, int nv, double *E) {
    REF_TRY
    Quiet q;
    cc_results r;
    r.T1 = t2(no, nv, T1), r.T2 = t4(no, no, nv, nv, T2);
    r.sliced_ints.oovv = t4(no, no, nv, nv, oovv), r.sliced_ints.vovv = t4(nv, no, nv, nv, vovv);
    r.sliced_ints.ovoo = t4(no, nv, no, no, ovoo);
    *E = CCSD_T(r, wrap_moes(foo, nullptr, fvv, no, nv));
    REF_CATCH
}

// The reference pipeline after the AO integrals: MP2() body (mp2.cpp:51-83) without eri_ao_tensor,
// then CCSD() (cc.cpp:409-451), make_moe_tensors + CCSD_T() (main.cpp:84-85).
// uhf = 0: C = Ca, eps = ea (Cb / eb ignored).  T1/T2/so_out may be NULL.  iters = last printed
// iteration number of the CCSD table (cc.cpp:437).
int ref_post_hf(const double *ao, const double *Ca, const double *Cb, const double *ea, const double *eb, int nao, int no,
                int nv, int uhf, int maxiter, double cc_conv, int do_t, double *e_mp2, double *e_ccsd, double *e_t,
                int *iters, double *T1_out, double *T2_out, double *T2_mp2_out) {
    REF_TRY
    Quiet q;
    params config;
    config.inputfile = "", config.type = do_t ? "CCSD(T)" : "CCSD", config.ref = uhf ? "UHF" : "RHF", config.basis = "";
    config.charge = 0, config.multiplicity = 1, config.maxiter = maxiter, config.scf_conv = cc_conv, config.cc_conv = cc_conv;
    scf_results scf;
    scf.energy = 0, scf.nalpha = scf.nbeta = 0, scf.no = no, scf.nv = nv, scf.nao = nao;
    mp2_results results;
    DTensor ao_ints = t4(nao, nao, nao, nao, ao);
    if (!uhf) {
        scf.C = mat(nao, nao, Ca), scf.moes = vec(nao, ea);
        auto mo_ints = transform_ao_mo(ao_ints, scf.C, scf.C);                     // mp2.cpp:55
        results.so_ints = transform_to_so(mo_ints, mo_ints, mo_ints);              // :58
        auto moes = make_so_moes(scf.moes, scf.moes, scf.nao);                     // :60
        auto oovv = slice_ints(results.so_ints, scf.no, scf.nv, "oovv");           // :62
        auto denom = make_denom(moes, scf.no, scf.nv);                             // :64
        results.energy = run_mp2(oovv, denom).energy;                              // :66
        results.T = run_mp2(oovv, denom).T;                                        // :67
    } else {
        scf.Ca = mat(nao, nao, Ca), scf.Cb = mat(nao, nao, Cb), scf.moes_a = vec(nao, ea), scf.moes_b = vec(nao, eb);
        auto mo_ints_aa = transform_ao_mo(ao_ints, scf.Ca, scf.Ca);                // :71
        auto mo_ints_bb = transform_ao_mo(ao_ints, scf.Cb, scf.Cb);                // :72
        auto mo_ints_ab = transform_ao_mo(ao_ints, scf.Ca, scf.Cb);                // :73
        results.so_ints = transform_to_so(mo_ints_aa, mo_ints_bb, mo_ints_ab);     // :75
        auto oovv = slice_ints(results.so_ints, scf.no, scf.nv, "oovv");           // :76
        auto moes = make_so_moes(scf.moes_a, scf.moes_b, scf.nao);                 // :78
        auto denom = make_denom(moes, scf.no, scf.nv);                             // :79
        results.energy = run_mp2(oovv, denom).energy;                              // :81
        results.T = run_mp2(oovv, denom).T;                                        // :82
    }
    *e_mp2 = results.energy;
    put(results.T, T2_mp2_out);
    if (e_ccsd) {
        std::fflush(stdout);
        cc_results cc = CCSD(scf, results, config);                                // main.cpp:83
        *e_ccsd = cc.ccsd_energy;
        put(cc.T1, T1_out), put(cc.T2, T2_out);
        if (do_t && e_t) {
            auto mo = make_moe_tensors(scf, config);                               // main.cpp:84
            *e_t = CCSD_T(cc, mo);                                                 // main.cpp:85
        }
        if (iters) *iters = count_cc_iterations(q.finish());
    }
    REF_CATCH
}

// number of basis functions of a molecule/basis (main.cpp:37-38), to size ref_molecule's outputs
int ref_nbasis(const char *xyz_path, const char *basis, int *nao_out) {
    REF_TRY
    Quiet q;
    std::vector<libint2::Atom> atoms = read_geometry(xyz_path);
    libint2::BasisSet obs(basis, atoms);
    *nao_out = (int) nbasis(obs.shells());
    REF_CATCH
}

// Whole reference program for one molecule, as main.cpp:22-88 runs it, with the energies returned
// at full precision (main prints 12 significant digits).  ref: "RHF" | "UHF"; type: "MP2" | "CCSD" | "CCSD(T)".
// Boundary tensors are handed out when the pointers are non-NULL (sizes from a first call with them NULL).
int ref_molecule(const char *xyz_path, const char *basis, const char *ref, const char *type, int multiplicity, int maxiter,
                 double scf_conv, double cc_conv, int *nao_out, int *no_out, int *nv_out, double *e_scf, double *e_mp2,
                 double *e_ccsd, double *e_t, int *cc_iters, double *ao_out, double *Ca_out, double *Cb_out, double *ea_out,
                 double *eb_out) {
    REF_TRY
    Quiet q;
    params config;
    config.inputfile = xyz_path, config.type = type, config.ref = ref, config.basis = basis;
    config.charge = 0, config.multiplicity = multiplicity, config.maxiter = maxiter, config.scf_conv = scf_conv, config.cc_conv = cc_conv;
    std::vector<libint2::Atom> atoms = read_geometry(config.inputfile);
    auto nelectron = 0;
    for (auto &atom : atoms) nelectron += atom.atomic_number;
    libint2::BasisSet obs(config.basis, atoms);
    const int nao = (int) nbasis(obs.shells());
    *nao_out = nao;
    const bool uhf = config.ref == "UHF";
    scf_results hf = uhf ? UHF(atoms, obs, nelectron, config) : RHF(atoms, obs, nelectron, config);
    *no_out = hf.no, *nv_out = hf.nv, *e_scf = hf.energy;
    const Matrix &A = uhf ? hf.Ca : hf.C, &B = uhf ? hf.Cb : hf.C;
    if (A.rows() != nao) throw std::runtime_error("reference SCF did not converge (results left unset, hf.cpp:445-454)");
    if (Ca_out) std::memcpy(Ca_out, A.data(), sizeof(double) * (size_t) nao * nao);
    if (Cb_out) std::memcpy(Cb_out, B.data(), sizeof(double) * (size_t) nao * nao);
    const Vector &va = uhf ? hf.moes_a : hf.moes, &vb = uhf ? hf.moes_b : hf.moes;
    for (int k = 0; k < nao; ++k) {
        if (ea_out) ea_out[k] = va(k);
        if (eb_out) eb_out[k] = vb(k);
    }
    if (ao_out) put(eri_ao_tensor(obs), ao_out);
    const std::string ty = type;
    if (ty == "MP2" || ty == "CCSD" || ty == "CCSD(T)") {
        auto mp2_result = MP2(obs, hf, config);
        *e_mp2 = mp2_result.energy;
        if (ty != "MP2") {
            std::fflush(stdout);
            auto ccsd_result = CCSD(hf, mp2_result, config);
            *e_ccsd = ccsd_result.ccsd_energy;
            if (ty == "CCSD(T)") {
                auto mo = make_moe_tensors(hf, config);
                *e_t = CCSD_T(ccsd_result, mo);
            }
        }
    }
    if (cc_iters) *cc_iters = count_cc_iterations(q.finish());
    REF_CATCH
}

}// extern "C"
