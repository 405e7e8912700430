// basis.cpp -- built-in basis-set data (STO-3G, cc-pVDZ, cc-pVTZ) and shell construction.
//
// Provenance: Libint's basis library is not in this image and there is no network, so the
// exponents / contraction coefficients below were typed in (Hehre-Stewart-Pople STO-3G;
// Dunning cc-pVDZ / cc-pVTZ, general contractions written out one contracted function per
// shell).  They are checked by (i) the contracted-norm self test (every contracted function of
// normalised primitives has norm 1 within 1e-3, tests/test_host_cpu.py) and (ii) the known
// answers quoted in tests/test_known_answers.py.  Energies depend only on the SPAN of the basis,
// so normalisation and function ordering conventions do not matter; a mistyped digit would
// shift total energies but cannot affect GPU-vs-CPU parity, which is asserted at the hot-path
// boundary on identical inputs.
#include "basis.h"

#include <algorithm>
#include <cctype>
#include <map>
#include <stdexcept>

namespace {

struct Contr {
    int l;
    std::vector<double> exps;
    std::vector<std::vector<double>> coefs;// one vector per contracted function (general contraction)
};
typedef std::vector<Contr> ElementBasis;

const std::vector<double> kSto1s{0.15432897, 0.53532814, 0.44463454};
const std::vector<double> kSto2s{-0.09996723, 0.39951283, 0.70011547};
const std::vector<double> kSto2p{0.15591627, 0.60768372, 0.39195739};

ElementBasis sto3g(int Z) {
    auto first_row = [](std::vector<double> e1, std::vector<double> e2) {
        return ElementBasis{{0, e1, {kSto1s}}, {0, e2, {kSto2s}}, {1, e2, {kSto2p}}};
    };
    switch (Z) {
        case 1: return ElementBasis{{0, {3.42525091, 0.62391373, 0.16885540}, {kSto1s}}};
        case 6: return first_row({71.6168370, 13.0450960, 3.5305122}, {2.9412494, 0.6834831, 0.2222899});
        case 7: return first_row({99.1061690, 18.0523120, 4.8856602}, {3.7804559, 0.8784966, 0.2857144});
        case 8: return first_row({130.7093200, 23.8088610, 6.4436083}, {5.0331513, 1.1695961, 0.3803890});
        default: throw std::invalid_argument("sto-3g: element not built in (H, C, N, O only)");
    }
}

ElementBasis ccpvdz(int Z) {
    auto unit = [](size_t n, size_t k) {
        std::vector<double> c(n, 0.0);
        c[k] = 1.0;
        return c;
    };
    switch (Z) {
        case 1:
            return ElementBasis{{0, {13.01, 1.962, 0.4446, 0.122}, {{0.019685, 0.137977, 0.478148, 0.501240}, unit(4, 3)}},
                                {1, {0.727}, {{1.0}}}};
        case 6:
            return ElementBasis{
                    {0, {6665.0, 1000.0, 228.0, 64.71, 21.06, 7.495, 2.797, 0.5215, 0.1596},
                     {{0.000692, 0.005329, 0.027077, 0.101718, 0.274740, 0.448564, 0.285074, 0.015204, -0.003191},
                      {-0.000146, -0.001154, -0.005725, -0.023312, -0.063955, -0.149981, -0.127262, 0.544529, 0.580496},
                      unit(9, 8)}},
                    {1, {9.439, 2.002, 0.5456, 0.1517}, {{0.038109, 0.209480, 0.508557, 0.468842}, unit(4, 3)}},
                    {2, {0.550}, {{1.0}}}};
        case 7:
            return ElementBasis{
                    {0, {9046.0, 1357.0, 309.3, 87.73, 28.56, 10.21, 3.838, 0.7466, 0.2248},
                     {{0.000700, 0.005389, 0.027406, 0.103207, 0.278723, 0.448540, 0.278238, 0.015440, -0.002864},
                      {-0.000153, -0.001208, -0.005992, -0.024544, -0.067459, -0.158078, -0.121831, 0.549003, 0.578815},
                      unit(9, 8)}},
                    {1, {13.55, 2.917, 0.7973, 0.2185}, {{0.039919, 0.217169, 0.510319, 0.462214}, unit(4, 3)}},
                    {2, {0.817}, {{1.0}}}};
        case 8:
            return ElementBasis{
                    {0, {11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013, 0.3023},
                     {{0.000710, 0.005470, 0.027837, 0.104800, 0.283062, 0.448719, 0.270952, 0.015458, -0.002585},
                      {-0.000160, -0.001263, -0.006267, -0.025716, -0.070924, -0.165411, -0.116955, 0.557368, 0.572759},
                      unit(9, 8)}},
                    {1, {17.70, 3.854, 1.046, 0.2753}, {{0.043018, 0.228913, 0.508728, 0.460531}, unit(4, 3)}},
                    {2, {1.185}, {{1.0}}}};
        default: throw std::invalid_argument("cc-pvdz: element not built in (H, C, N, O only)");
    }
}

ElementBasis ccpvtz(int Z) {
    auto unit = [](size_t n, size_t k) {
        std::vector<double> c(n, 0.0);
        c[k] = 1.0;
        return c;
    };
    switch (Z) {
        case 1:
            return ElementBasis{{0, {33.87, 5.095, 1.159, 0.3258, 0.1027},
                                 {{0.006068, 0.045308, 0.202822, 0.503903, 0.383421}, unit(5, 3), unit(5, 4)}},
                                {1, {1.407, 0.388}, {unit(2, 0), unit(2, 1)}},
                                {2, {1.057}, {{1.0}}}};
        case 6:
            return ElementBasis{
                    {0, {8236.0, 1235.0, 280.8, 79.27, 25.59, 8.997, 3.319, 0.9059, 0.3643, 0.1285},
                     {{0.000531, 0.004108, 0.021087, 0.081853, 0.234817, 0.434401, 0.346129, 0.039378, -0.008983, 0.002385},
                      {-0.000113, -0.000878, -0.004540, -0.018133, -0.055760, -0.126895, -0.170352, 0.140382, 0.598684, 0.395389},
                      unit(10, 7), unit(10, 9)}},
                    {1, {18.71, 4.133, 1.200, 0.3827, 0.1209},
                     {{0.014031, 0.086866, 0.290216, 0.501008, 0.343406}, unit(5, 3), unit(5, 4)}},
                    {2, {1.097, 0.318}, {unit(2, 0), unit(2, 1)}},
                    {3, {0.761}, {{1.0}}}};
        default: throw std::invalid_argument("cc-pvtz: element not built in (H, C only)");
    }
}

}// namespace

BasisSet::BasisSet(const std::string &name, const std::vector<Atom> &atoms) {
    std::string key = name;
    std::transform(key.begin(), key.end(), key.begin(), [](unsigned char c) { return std::tolower(c); });
    for (const Atom &a : atoms) {
        ElementBasis eb;
        if (key == "sto-3g") eb = sto3g(a.atomic_number);
        else if (key == "cc-pvdz") eb = ccpvdz(a.atomic_number);
        else if (key == "cc-pvtz") eb = ccpvtz(a.atomic_number);
        else throw std::invalid_argument("basis set " + name + " is not built in (sto-3g, cc-pvdz, cc-pvtz)");
        for (const Contr &c : eb)
            for (const auto &coef : c.coefs) {
                Shell s;
                s.l = c.l;
                s.O = {{a.x, a.y, a.z}};
                for (size_t k = 0; k < c.exps.size(); ++k)
                    if (coef[k] != 0.0) s.alpha.push_back(c.exps[k]), s.coef.push_back(coef[k]);
                shells_.push_back(s);
            }
    }
}

std::vector<size_t> BasisSet::shell2bf() const {
    std::vector<size_t> r;
    size_t n = 0;
    for (const Shell &s : shells_) r.push_back(n), n += s.size();
    return r;
}
size_t BasisSet::nbf() const { return nbasis(shells_); }
int BasisSet::max_l() const {
    int l = 0;
    for (const Shell &s : shells_) l = std::max(l, s.l);
    return l;
}
size_t nbasis(const std::vector<Shell> &shells) {
    size_t n = 0;
    for (const Shell &s : shells) n += s.size();
    return n;
}
