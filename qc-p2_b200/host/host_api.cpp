// host_api.cpp -- flat C entry points of the HOST substrate for the Python tests and tools:
// run the SCF of a molecule and hand out the hot-path boundary inputs (AO integrals, MO
// coefficients, orbital energies).  No GPU code is reachable from here.
#include <cstring>
#include <iostream>
#include <sstream>

#include "hf.h"

extern bool g_scf_quiet;

extern "C" {

// xyz_path: geometry in Angstrom; basis: "sto-3g" | "cc-pvdz" | "cc-pvtz"; uhf: 0/1.
// Outputs (caller-allocated, n = *nao_out obtained from a first call with ao == NULL):
//   ao[n^4], Ca[n^2], Cb[n^2] (RHF: Cb = Ca), ea[n], eb[n], scalars.  Returns 0 on success.
int qcp2_host_scf(const char *xyz_path, const char *basis, int uhf, int multiplicity, int maxiter, double conv, int *nao_out,
                  int *no_out, int *nv_out, double *e_scf, double *e_nuc, int *iters, double *ao, double *Ca, double *Cb,
                  double *ea, double *eb, double *S_out, int quiet) {
    std::streambuf *old = nullptr;
    std::ostringstream sink;
    if (quiet) old = std::cout.rdbuf(sink.rdbuf());
    g_scf_quiet = quiet != 0;
    int rc = 0;
    try {
        std::vector<Atom> atoms = read_geometry(xyz_path);
        BasisSet obs(basis, atoms);
        const int n = (int) obs.nbf();
        *nao_out = n;
        int nel = 0;
        for (auto &a : atoms) nel += a.atomic_number;
        if (ao) {
            params cfg;
            cfg.inputfile = xyz_path, cfg.type = "CCSD(T)", cfg.ref = uhf ? "UHF" : "RHF", cfg.basis = basis;
            cfg.charge = 0, cfg.multiplicity = multiplicity, cfg.maxiter = maxiter, cfg.scf_conv = conv, cfg.cc_conv = conv;
            DTensor eri = eri_ao_tensor(obs);
            scf_results r = uhf ? UHF(atoms, obs, nel, cfg, &eri) : RHF(atoms, obs, nel, cfg, &eri);
            if (!r.converged) throw std::runtime_error("SCF did not converge");
            std::memcpy(ao, eri.data(), sizeof(double) * eri.size());
            const Matrix &A = uhf ? r.Ca : r.C, &B = uhf ? r.Cb : r.C;
            std::memcpy(Ca, A.data(), sizeof(double) * (size_t) n * n);
            std::memcpy(Cb, B.data(), sizeof(double) * (size_t) n * n);
            const Vector &va = uhf ? r.moes_a : r.moes, &vb = uhf ? r.moes_b : r.moes;
            std::memcpy(ea, va.data(), sizeof(double) * (size_t) n);
            std::memcpy(eb, vb.data(), sizeof(double) * (size_t) n);
            *no_out = r.no, *nv_out = r.nv, *e_scf = r.energy, *e_nuc = compute_enuc(atoms), *iters = r.iterations;
            if (S_out) {
                Matrix S = compute_1body_ints(obs, Operator1::overlap);
                std::memcpy(S_out, S.data(), sizeof(double) * (size_t) n * n);
            }
        }
    } catch (const std::exception &e) {
        std::cerr << "qcp2_host_scf: " << e.what() << std::endl;
        rc = 1;
    }
    if (quiet) std::cout.rdbuf(old);
    return rc;
}

// contracted norms of every shell of a basis on a single atom (data self-test)
int qcp2_host_basis_norms(const char *basis, int Z, double *norms, int *ls, int max_shells) {
    try {
        BasisSet obs(basis, std::vector<Atom>{Atom{Z, 0, 0, 0}});
        int k = 0;
        for (const Shell &s : obs.shells()) {
            if (k >= max_shells) break;
            norms[k] = contracted_norm(s), ls[k] = s.l, ++k;
        }
        return k;
    } catch (const std::exception &e) {
        std::cerr << "qcp2_host_basis_norms: " << e.what() << std::endl;
        return -1;
    }
}
}
