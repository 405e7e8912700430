// hf.h -- host SCF (RHF / UHF), same iteration scheme and result struct as the reference's
// src/hf.{h,cpp}.  OUTSIDE the hot path (north_star: "Libint AO integral generation and the
// small SCF diagonalisation stay on the host").
#pragma once
#include "basis.h"
#include "gaussian.h"
#include "general.h"

struct scf_results {// hf.h:29-35
    real_t energy = 0.0;
    int nalpha = 0, nbeta = 0, no = 0, nv = 0;
    int nao = 0;
    Matrix F, Fa, Fb, C, Ca, Cb, D, Da, Db;
    Vector moes, moes_a, moes_b;
    bool converged = false;
    int iterations = 0;
};

Matrix density_guess(int nocc, int nao);                                                     // hf.cpp:116-121
// Fock two-electron part from the stored AO tensor; same G as build_fock / build_uhf_fock
// (hf.cpp:124-241, 243-324), which recompute the integrals every iteration instead
Matrix build_fock(const DTensor &ao_ints, const Matrix &D);
Matrix build_uhf_fock(const DTensor &ao_ints, const Matrix &D, const Matrix &Ds);
real_t rhf_energy(const Matrix &D, const Matrix &H, const Matrix &F);                        // hf.cpp:327-333
real_t uhf_energy(const Matrix &D, const Matrix &Da, const Matrix &Db, const Matrix &H, const Matrix &Fa,
                  const Matrix &Fb);                                                         // hf.cpp:335-342
// F C = e S C, eigenvalues ascending (Eigen::GeneralizedSelfAdjointEigenSolver in the reference)
void generalized_eigen(const Matrix &F, const Matrix &S, Vector &eps, Matrix &C);

scf_results RHF(const std::vector<Atom> &atoms, const BasisSet &obs, real_t nelectron, params config,
                const DTensor *ao_ints = nullptr);
scf_results UHF(const std::vector<Atom> &atoms, const BasisSet &obs, real_t nelectron, params config,
                const DTensor *ao_ints = nullptr);
