// libint2_shim.cpp -- TEST INFRASTRUCTURE ONLY (oracle/ref_shim): implementation of the
// libint2 stand-in declared in include/libint2.hpp on top of backend.h.
#include <libint2.hpp>

#include <algorithm>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>

#include "backend.h"

namespace libint2 {

struct BackendBasis {
    void *h = nullptr;
    std::vector<Shell> shells;
    std::vector<size_t> first;// shell -> first basis function
    size_t nbf = 0;
    std::vector<double> S, T, V, eri;
    std::vector<std::pair<double, std::array<double, 3>>> V_charges;
    ~BackendBasis() {
        if (h) shimb_basis_destroy(h);
    }
    int find(const Shell &s) const {
        for (size_t k = 0; k < shells.size(); ++k)
            if (shells[k] == s) return static_cast<int>(k);
        return -1;
    }
};

namespace {
std::vector<std::weak_ptr<BackendBasis>> &registry() {
    static std::vector<std::weak_ptr<BackendBasis>> r;
    return r;
}
// most recently created basis that contains every one of the given shells
std::shared_ptr<BackendBasis> lookup(std::initializer_list<const Shell *> sh, int *idx) {
    auto &r = registry();
    for (size_t k = r.size(); k-- > 0;) {
        auto b = r[k].lock();
        if (!b) continue;
        bool ok = true;
        int n = 0;
        for (const Shell *s : sh) {
            const int i = b->find(*s);
            if (i < 0) {
                ok = false;
                break;
            }
            idx[n++] = i;
        }
        if (ok) return b;
    }
    throw std::runtime_error("libint2 shim: Engine::compute on shells that belong to no BasisSet(name, atoms)");
}
}// namespace

BasisSet::BasisSet(const std::string &name, const std::vector<Atom> &atoms) {
    std::vector<int> Z;
    std::vector<double> xyz;
    for (const Atom &a : atoms) Z.push_back(a.atomic_number), xyz.push_back(a.x), xyz.push_back(a.y), xyz.push_back(a.z);
    auto b = std::make_shared<BackendBasis>();
    b->h = shimb_basis_create(name.c_str(), static_cast<int>(atoms.size()), Z.data(), xyz.data());
    if (!b->h) throw std::invalid_argument("libint2 shim: basis set " + name + " not available");
    const int ns = shimb_nshell(b->h);
    size_t off = 0;
    for (int s = 0; s < ns; ++s) {
        Shell sh;
        double al[64], co[64], O[3];
        const int np = shimb_shell(b->h, s, &sh.l, O, al, co, 64);
        sh.pure = sh.l >= 2;
        sh.O = {{O[0], O[1], O[2]}};
        sh.alpha.assign(al, al + np), sh.coeff.assign(co, co + np);
        b->first.push_back(off);
        off += sh.size();
        b->shells.push_back(sh);
    }
    b->nbf = off;
    if (static_cast<int>(off) != shimb_nbf(b->h)) throw std::runtime_error("libint2 shim: basis-function count mismatch");
    shells_ = b->shells;
    backend_ = b;
    registry().push_back(b);
}

BasisSet::BasisSet(const std::vector<Shell> &shells) : shells_(shells) {
    auto &r = registry();
    for (size_t k = r.size(); k-- > 0;) {
        auto b = r[k].lock();
        if (b && b->shells == shells) {
            backend_ = b;
            return;
        }
    }
    throw std::runtime_error("libint2 shim: BasisSet(vector<Shell>) from shells of no known basis");
}

size_t BasisSet::max_nprim() const {
    size_t n = 0;
    for (const Shell &s : shells_) n = std::max(n, s.alpha.size());
    return n;
}
int BasisSet::max_l() const {
    int l = 0;
    for (const Shell &s : shells_) l = std::max(l, s.l);
    return l;
}
std::vector<size_t> BasisSet::shell2bf() const {
    std::vector<size_t> r;
    size_t n = 0;
    for (const Shell &s : shells_) r.push_back(n), n += s.size();
    return r;
}

Engine::Engine(Operator op, size_t, int, int deriv) : op_(op), results_(1, nullptr) {
    if (deriv != 0) throw std::invalid_argument("libint2 shim: derivative integrals are not provided");
}

void Engine::compute(const Shell &s1, const Shell &s2) {
    int ix[2];
    auto b = lookup({&s1, &s2}, ix);
    const size_t n = b->nbf;
    std::vector<double> *M = nullptr;
    int opcode = 0;
    switch (op_) {
        case Operator::overlap: M = &b->S, opcode = 0; break;
        case Operator::kinetic: M = &b->T, opcode = 1; break;
        case Operator::nuclear: M = &b->V, opcode = 2; break;
        default: throw std::invalid_argument("libint2 shim: two-centre compute with a two-electron operator");
    }
    if (op_ == Operator::nuclear && b->V_charges != charges_) M->clear();
    if (M->empty()) {
        std::vector<int> Z;
        std::vector<double> xyz;
        for (const auto &q : charges_) {
            Z.push_back(static_cast<int>(std::lround(q.first)));
            xyz.push_back(q.second[0]), xyz.push_back(q.second[1]), xyz.push_back(q.second[2]);
        }
        M->resize(n * n);
        if (shimb_onebody(b->h, opcode, static_cast<int>(Z.size()), Z.data(), xyz.data(), M->data()))
            throw std::runtime_error("libint2 shim: one-body integrals failed");
        if (op_ == Operator::nuclear) b->V_charges = charges_;
    }
    const size_t n1 = s1.size(), n2 = s2.size(), f1 = b->first[static_cast<size_t>(ix[0])], f2 = b->first[static_cast<size_t>(ix[1])];
    scratch_.resize(n1 * n2);
    for (size_t a = 0; a < n1; ++a)
        for (size_t c = 0; c < n2; ++c) scratch_[a * n2 + c] = (*M)[(f1 + a) * n + f2 + c];
    results_[0] = scratch_.data();
}

void Engine::compute(const Shell &s1, const Shell &s2, const Shell &s3, const Shell &s4) {
    if (op_ != Operator::coulomb) throw std::invalid_argument("libint2 shim: four-centre compute needs Operator::coulomb");
    int ix[4];
    auto b = lookup({&s1, &s2, &s3, &s4}, ix);
    const size_t n = b->nbf;
    if (b->eri.empty()) {
        b->eri.resize(n * n * n * n);
        if (shimb_eri(b->h, b->eri.data())) throw std::runtime_error("libint2 shim: two-electron integrals failed");
    }
    const size_t n1 = s1.size(), n2 = s2.size(), n3 = s3.size(), n4 = s4.size();
    const size_t f1 = b->first[static_cast<size_t>(ix[0])], f2 = b->first[static_cast<size_t>(ix[1])],
                 f3 = b->first[static_cast<size_t>(ix[2])], f4 = b->first[static_cast<size_t>(ix[3])];
    scratch_.resize(n1 * n2 * n3 * n4);
    size_t k = 0;
    for (size_t a = 0; a < n1; ++a)
        for (size_t c = 0; c < n2; ++c)
            for (size_t d = 0; d < n3; ++d) {
                const double *src = &b->eri[(((f1 + a) * n + f2 + c) * n + f3 + d) * n + f4];
                for (size_t e = 0; e < n4; ++e) scratch_[k++] = src[e];
            }
    results_[0] = scratch_.data();
}

std::vector<Atom> read_dotxyz(std::istream &is) {
    static const char *names[] = {"X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar"};
    // CODATA 2018 bohr radius, what current Libint's read_dotxyz defaults to; the same constant as
    // the product's host reader (qc-p2_b200/host/general.cpp), so both see identical geometries
    const double ang_to_bohr = 1.0 / 0.529177210903;
    std::string line;
    std::getline(is, line);
    const int natom = std::atoi(line.c_str());
    std::getline(is, line);// comment
    std::vector<Atom> atoms;
    for (int k = 0; k < natom; ++k) {
        if (!std::getline(is, line)) throw std::invalid_argument("libint2 shim: xyz file ended early");
        std::istringstream ls(line);
        std::string sym;
        double x, y, z;
        if (!(ls >> sym >> x >> y >> z)) throw std::invalid_argument("libint2 shim: malformed xyz line: " + line);
        if (!sym.empty()) sym[0] = static_cast<char>(std::toupper(static_cast<unsigned char>(sym[0])));
        for (size_t c = 1; c < sym.size(); ++c) sym[c] = static_cast<char>(std::tolower(static_cast<unsigned char>(sym[c])));
        int Z = 0;
        for (int q = 1; q < 19; ++q)
            if (sym == names[q]) Z = q;
        if (!Z) throw std::invalid_argument("libint2 shim: unknown element " + sym);
        atoms.push_back(Atom{Z, x * ang_to_bohr, y * ang_to_bohr, z * ang_to_bohr});
    }
    return atoms;
}

size_t sto3g_num_ao(size_t Z) {
    if (Z <= 2) return 1;
    if (Z <= 10) return 5;
    if (Z <= 18) return 9;
    throw std::invalid_argument("libint2 shim: sto3g_num_ao only for Z <= 18");
}

// electrons smeared evenly over the orbitals of each subshell, filled in aufbau order
const std::vector<double> &sto3g_ao_occupation_vector(size_t Z) {
    static std::map<size_t, std::vector<double>> cache;
    auto it = cache.find(Z);
    if (it != cache.end()) return it->second;
    std::vector<double> occ(sto3g_num_ao(Z), 0.0);
    double left = static_cast<double>(Z);
    const size_t sub[] = {1, 1, 3, 1, 3};// 1s 2s 2p 3s 3p
    size_t ao = 0;
    for (size_t s = 0; s < 5 && ao < occ.size(); ++s) {
        const double cap = 2.0 * static_cast<double>(sub[s]);
        const double e = std::min(left, cap);
        for (size_t k = 0; k < sub[s]; ++k) occ[ao + k] = e / static_cast<double>(sub[s]);
        ao += sub[s], left -= e;
    }
    return cache.emplace(Z, occ).first->second;
}

}// namespace libint2
