// libint2.hpp -- TEST INFRASTRUCTURE ONLY (oracle/ref_shim).
//
// Stand-in for the Libint2 surface the reference uses: Atom, Shell,
// BasisSet, Engine, Operator, read_dotxyz, initialize/finalize, scalar_type.  Libint2 and its
// basis-set library are not in this image.  The numbers come from the repo's own host substrate
// (qc-p2_b200/host/{basis,gaussian}.cpp, reached through the flat C interface in backend.h), so
// the reference sources compiled against this header see exactly the AO integrals the product's
// host side produces -- i.e. identical hot-path boundary inputs.  The post-HF hot path itself
// (integrals.cpp:95-207, mp2.cpp, cc.cpp) touches none of this except eri_ao_tensor's loop.
#pragma once
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <fstream>
#include <iostream>
#include <istream>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include <math.h>
#include <stdlib.h>

#include <Eigen/Eigenvalues>

#define LIBINT2_CONSTEXPR_STATICS 1

namespace libint2 {

typedef double scalar_type;
typedef double real_t;

struct Atom {
    int atomic_number;
    double x, y, z;// bohr
};

struct Shell {
    int l = 0;
    bool pure = false;
    std::vector<double> alpha, coeff;
    std::array<double, 3> O{{0, 0, 0}};
    size_t size() const { return pure ? static_cast<size_t>(2 * l + 1) : static_cast<size_t>((l + 1) * (l + 2) / 2); }
    bool operator==(const Shell &o) const { return l == o.l && pure == o.pure && O == o.O && alpha == o.alpha && coeff == o.coeff; }
};

enum class Operator { overlap, kinetic, nuclear, coulomb };

struct BackendBasis;// shim-internal: backend handle + cached integral matrices

class BasisSet {
public:
    BasisSet() {}
    BasisSet(const std::string &name, const std::vector<Atom> &atoms);
    BasisSet(const std::vector<Shell> &shells);// hf.cpp:419,529-531 pass obs.shells() where a BasisSet is expected
    const std::vector<Shell> &shells() const { return shells_; }
    size_t max_nprim() const;
    int max_l() const;
    std::vector<size_t> shell2bf() const;
    const std::shared_ptr<BackendBasis> &backend() const { return backend_; }

private:
    std::vector<Shell> shells_;
    std::shared_ptr<BackendBasis> backend_;
};

class Engine {
public:
    typedef std::vector<const double *> target_ptr_vec;
    Engine(Operator op, size_t max_nprim, int max_l, int deriv_order = 0);
    void set_params(const std::vector<std::pair<double, std::array<double, 3>>> &charges) { charges_ = charges; }
    const target_ptr_vec &results() const { return results_; }
    void compute(const Shell &s1, const Shell &s2);
    void compute(const Shell &s1, const Shell &s2, const Shell &s3, const Shell &s4);

private:
    Operator op_;
    std::vector<std::pair<double, std::array<double, 3>>> charges_;
    std::vector<double> scratch_;
    target_ptr_vec results_;
};

inline void initialize() {}
inline void finalize() {}

// XYZ reader: atom count, comment, "symbol x y z" in Angstrom -> bohr (general.cpp:65)
std::vector<Atom> read_dotxyz(std::istream &is);

// libint2/chemistry/sto3g_atomic_density.h (hf.cpp:96-104; dead unless basis == "STO-3G" exactly)
size_t sto3g_num_ao(size_t Z);
const std::vector<double> &sto3g_ao_occupation_vector(size_t Z);

}// namespace libint2
