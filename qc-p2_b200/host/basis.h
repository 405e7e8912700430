// basis.h -- Gaussian basis sets for the host SCF substrate (stand-in for libint2::BasisSet,
// main.cpp:37).  Libint and its basis-set library are not available in this image, so the
// data for the bases named by BASELINE.json's configs (sto-3g, cc-pvdz, cc-pvtz; H, C, N, O)
// is built in (basis.cpp).  This is OUTSIDE the hot path: it only produces the boundary
// inputs (AO integrals, SCF coefficients, orbital energies) for the real-molecule configs.
#pragma once
#include <array>
#include <string>
#include <vector>

#include "general.h"

struct Shell {
    int l = 0;                      // angular momentum; l >= 2 shells are pure (2l+1 functions)
    std::vector<double> alpha;      // primitive exponents
    std::vector<double> coef;       // contraction coefficients of normalised primitives
    std::array<double, 3> O{{0, 0, 0}};
    size_t size() const { return l < 2 ? (size_t) ((l + 1) * (l + 2) / 2) : (size_t) (2 * l + 1); }
    int ncart() const { return (l + 1) * (l + 2) / 2; }
};

class BasisSet {
public:
    BasisSet() {}
    // name is matched case-insensitively ("sto-3g", "cc-pvdz", "cc-pvtz")
    BasisSet(const std::string &name, const std::vector<Atom> &atoms);
    const std::vector<Shell> &shells() const { return shells_; }
    std::vector<size_t> shell2bf() const;
    size_t nbf() const;
    int max_l() const;

private:
    std::vector<Shell> shells_;
};

size_t nbasis(const std::vector<Shell> &shells);// general.cpp:82-87
