// gaussian.cpp -- one- and two-electron integrals over contracted Gaussians (McMurchie-
// Davidson), the host-side replacement for the Libint engines the reference calls in
// hf.cpp:37-88 and integrals.cpp:12-89.  s, p cartesian; d, f pure (solid harmonics), as in
// Libint's cc-pVXZ library.  OUTSIDE the hot path: it only builds the boundary inputs.
#include "gaussian.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace {

const double kPi = 3.14159265358979323846;

// ---- Boys function F_0..F_nmax(T) --------------------------------------------------
void boys(int nmax, double T, double *F) {
    if (T < 35.0) {
        // series for the highest order, then stable downward recursion
        double term = 1.0 / (2 * nmax + 1), sum = term;
        for (int k = 1; k < 400; ++k) {
            term *= 2.0 * T / (2 * nmax + 2 * k + 1);
            sum += term;
            if (term < 1e-17 * sum) break;
        }
        const double eT = std::exp(-T);
        F[nmax] = eT * sum;
        for (int n = nmax; n > 0; --n) F[n - 1] = (2.0 * T * F[n] + eT) / (2 * n - 1);
    } else {
        // erfc(sqrt(T)) < 1e-16 here: asymptotic F_0, upward recursion is stable for large T
        const double eT = std::exp(-T);
        F[0] = 0.5 * std::sqrt(kPi / T) * std::erf(std::sqrt(T));
        for (int n = 0; n < nmax; ++n) F[n + 1] = ((2 * n + 1) * F[n] - eT) / (2.0 * T);
    }
}

// ---- Hermite expansion coefficients E_t^{ij} for one Cartesian direction ---------------
// E[(i*(jmax+1)+j)*(imax+jmax+1)+t]
struct ETable {
    int imax, jmax, tdim;
    std::vector<double> e;
    double operator()(int i, int j, int t) const {
        if (t < 0 || t > i + j) return 0.0;
        return e[(size_t) ((i * (jmax + 1) + j) * tdim + t)];
    }
};
ETable hermite_E(int imax, int jmax, double a, double b, double Qx) {
    ETable T;
    T.imax = imax, T.jmax = jmax, T.tdim = imax + jmax + 1;
    T.e.assign((size_t) ((imax + 1) * (jmax + 1) * T.tdim), 0.0);
    const double p = a + b, q = a * b / p;
    auto at = [&](int i, int j, int t) -> double & { return T.e[(size_t) ((i * (jmax + 1) + j) * T.tdim + t)]; };
    auto get = [&](int i, int j, int t) -> double { return (t < 0 || t > i + j) ? 0.0 : at(i, j, t); };
    at(0, 0, 0) = std::exp(-q * Qx * Qx);
    for (int i = 1; i <= imax; ++i)
        for (int t = 0; t <= i; ++t)
            at(i, 0, t) = get(i - 1, 0, t - 1) / (2 * p) - (q * Qx / a) * get(i - 1, 0, t) + (t + 1) * get(i - 1, 0, t + 1);
    for (int i = 0; i <= imax; ++i)
        for (int j = 1; j <= jmax; ++j)
            for (int t = 0; t <= i + j; ++t)
                at(i, j, t) = get(i, j - 1, t - 1) / (2 * p) + (q * Qx / b) * get(i, j - 1, t) + (t + 1) * get(i, j - 1, t + 1);
    return T;
}

// ---- Hermite Coulomb integrals R_{tuv}, t+u+v <= L --------------------------------------
struct RTable {
    int L;
    std::vector<double> r;// [t][u][v], dim (L+1)^3
    double operator()(int t, int u, int v) const { return r[(size_t) ((t * (L + 1) + u) * (L + 1) + v)]; }
};
void hermite_R(int L, double p, const double PC[3], RTable &out, std::vector<double> &work) {
    const int D = L + 1;
    out.L = L;
    out.r.assign((size_t) (D * D * D), 0.0);
    work.assign((size_t) (D * D * D * D), 0.0);// [n][t][u][v]
    auto W = [&](int n, int t, int u, int v) -> double & { return work[(size_t) (((n * D + t) * D + u) * D + v)]; };
    double F[32];
    boys(L, p * (PC[0] * PC[0] + PC[1] * PC[1] + PC[2] * PC[2]), F);
    double f = 1.0;
    for (int n = 0; n <= L; ++n) W(n, 0, 0, 0) = f * F[n], f *= -2.0 * p;
    for (int tot = 1; tot <= L; ++tot)
        for (int t = 0; t <= tot; ++t)
            for (int u = 0; u <= tot - t; ++u) {
                const int v = tot - t - u;
                for (int n = 0; n <= L - tot; ++n) {
                    double val;
                    if (t > 0) val = (t > 1 ? (t - 1) * W(n + 1, t - 2, u, v) : 0.0) + PC[0] * W(n + 1, t - 1, u, v);
                    else if (u > 0) val = (u > 1 ? (u - 1) * W(n + 1, t, u - 2, v) : 0.0) + PC[1] * W(n + 1, t, u - 1, v);
                    else val = (v > 1 ? (v - 1) * W(n + 1, t, u, v - 2) : 0.0) + PC[2] * W(n + 1, t, u, v - 1);
                    W(n, t, u, v) = val;
                }
            }
    for (int t = 0; t <= L; ++t)
        for (int u = 0; u <= L - t; ++u)
            for (int v = 0; v <= L - t - u; ++v) out.r[(size_t) ((t * D + u) * D + v)] = W(0, t, u, v);
}

// ---- Cartesian components and pure transformations ------------------------------------
struct Cart {
    int x, y, z;
};
std::vector<Cart> cart_components(int l) {
    std::vector<Cart> c;
    for (int x = l; x >= 0; --x)
        for (int y = l - x; y >= 0; --y) c.push_back(Cart{x, y, l - x - y});
    return c;
}
int cart_index(int l, int x, int y) {// position of (x,y,l-x-y) in cart_components(l)
    int idx = 0;
    for (int xx = l; xx > x; --xx) idx += l - xx + 1;
    return idx + (l - x - y);
}
double dfact(int n) { return n <= 1 ? 1.0 : n * dfact(n - 2); }
double prim_norm(double a, int l) {// norm of x^l exp(-a r^2)
    return std::pow(2.0 * a / kPi, 0.75) * std::pow(4.0 * a, 0.5 * l) / std::sqrt(dfact(2 * l - 1));
}
// rows: real solid harmonics as polynomials in Cartesian monomials (unnormalised)
std::vector<std::vector<double>> raw_transform(int l) {
    const int nc = (l + 1) * (l + 2) / 2;
    auto row = [&](std::initializer_list<std::pair<std::pair<int, int>, double>> terms) {
        std::vector<double> r((size_t) nc, 0.0);
        for (auto &t : terms) r[(size_t) cart_index(l, t.first.first, t.first.second)] = t.second;
        return r;
    };
    auto m = [](int x, int y) { return std::make_pair(x, y); };
    std::vector<std::vector<double>> T;
    if (l < 2) {
        for (int k = 0; k < nc; ++k) {
            std::vector<double> r((size_t) nc, 0.0);
            r[(size_t) k] = 1.0;
            T.push_back(r);
        }
    } else if (l == 2) {
        T.push_back(row({{m(1, 1), 1.0}}));                                   // xy
        T.push_back(row({{m(0, 1), 1.0}}));                                   // yz
        T.push_back(row({{m(0, 0), 2.0}, {m(2, 0), -1.0}, {m(0, 2), -1.0}})); // 2zz-xx-yy
        T.push_back(row({{m(1, 0), 1.0}}));                                   // xz
        T.push_back(row({{m(2, 0), 1.0}, {m(0, 2), -1.0}}));                  // xx-yy
    } else if (l == 3) {
        T.push_back(row({{m(2, 1), 3.0}, {m(0, 3), -1.0}}));                  // 3xxy - yyy
        T.push_back(row({{m(1, 1), 1.0}}));                                   // xyz
        T.push_back(row({{m(0, 1), 4.0}, {m(2, 1), -1.0}, {m(0, 3), -1.0}})); // y(4zz-xx-yy)
        T.push_back(row({{m(0, 0), 2.0}, {m(2, 0), -3.0}, {m(0, 2), -3.0}})); // z(2zz-3xx-3yy)
        T.push_back(row({{m(1, 0), 4.0}, {m(3, 0), -1.0}, {m(1, 2), -1.0}})); // x(4zz-xx-yy)
        T.push_back(row({{m(2, 0), 1.0}, {m(0, 2), -1.0}}));                  // z(xx-yy)
        T.push_back(row({{m(3, 0), 1.0}, {m(1, 2), -3.0}}));                  // xxx - 3xyy
    } else {
        throw std::invalid_argument("angular momentum > 3 is not supported by the host substrate");
    }
    return T;
}

// Cartesian overlap block of two shells (raw monomials, primitive norms folded into weights)
void overlap_cart(const Shell &A, const Shell &B, std::vector<double> &S) {
    const auto ca = cart_components(A.l), cb = cart_components(B.l);
    S.assign(ca.size() * cb.size(), 0.0);
    for (size_t i = 0; i < A.alpha.size(); ++i)
        for (size_t j = 0; j < B.alpha.size(); ++j) {
            const double a = A.alpha[i], b = B.alpha[j], p = a + b;
            const double w = A.coef[i] * prim_norm(a, A.l) * B.coef[j] * prim_norm(b, B.l) * std::pow(kPi / p, 1.5);
            ETable E[3];
            for (int d = 0; d < 3; ++d) E[d] = hermite_E(A.l, B.l, a, b, A.O[(size_t) d] - B.O[(size_t) d]);
            for (size_t x = 0; x < ca.size(); ++x)
                for (size_t y = 0; y < cb.size(); ++y)
                    S[x * cb.size() + y] += w * E[0](ca[x].x, cb[y].x, 0) * E[1](ca[x].y, cb[y].y, 0) * E[2](ca[x].z, cb[y].z, 0);
        }
}

// per-shell transformation (size() x ncart()) normalised to unit self-overlap
std::vector<std::vector<double>> shell_transform(const Shell &s) {
    auto T = raw_transform(s.l);
    std::vector<double> S;
    overlap_cart(s, s, S);
    const size_t nc = (size_t) s.ncart();
    for (auto &r : T) {
        double nn = 0.0;
        for (size_t x = 0; x < nc; ++x)
            for (size_t y = 0; y < nc; ++y) nn += r[x] * S[x * nc + y] * r[y];
        const double f = 1.0 / std::sqrt(nn);
        for (double &c : r) c *= f;
    }
    return T;
}

// out[pa][pb] = sum Ta[pa][ca] in[ca][cb] Tb[pb][cb]
void transform2(const std::vector<std::vector<double>> &Ta, const std::vector<std::vector<double>> &Tb,
                const std::vector<double> &in, size_t nca, size_t ncb, std::vector<double> &out) {
    out.assign(Ta.size() * Tb.size(), 0.0);
    for (size_t pa = 0; pa < Ta.size(); ++pa)
        for (size_t x = 0; x < nca; ++x) {
            if (Ta[pa][x] == 0.0) continue;
            for (size_t pb = 0; pb < Tb.size(); ++pb)
                for (size_t y = 0; y < ncb; ++y) out[pa * Tb.size() + pb] += Ta[pa][x] * Tb[pb][y] * in[x * ncb + y];
        }
}

struct ShellData {
    std::vector<Cart> cart;
    std::vector<std::vector<double>> T;
};

// Hermite expansion of one primitive pair for every Cartesian component pair
struct PrimPair {
    double p, P[3], w;
    int L;
    std::vector<double> E;// [ia][ib][t][u][v], (L+1)^3 per component pair
};
struct ShellPair {
    std::vector<PrimPair> prims;
};

ShellPair make_pair_data(const Shell &A, const Shell &B) {
    ShellPair sp;
    const auto ca = cart_components(A.l), cb = cart_components(B.l);
    const int L = A.l + B.l, D = L + 1;
    for (size_t i = 0; i < A.alpha.size(); ++i)
        for (size_t j = 0; j < B.alpha.size(); ++j) {
            PrimPair pp;
            const double a = A.alpha[i], b = B.alpha[j];
            pp.p = a + b, pp.L = L;
            for (int d = 0; d < 3; ++d) pp.P[d] = (a * A.O[(size_t) d] + b * B.O[(size_t) d]) / pp.p;
            pp.w = A.coef[i] * prim_norm(a, A.l) * B.coef[j] * prim_norm(b, B.l);
            ETable E[3];
            for (int d = 0; d < 3; ++d) E[d] = hermite_E(A.l, B.l, a, b, A.O[(size_t) d] - B.O[(size_t) d]);
            pp.E.assign(ca.size() * cb.size() * (size_t) (D * D * D), 0.0);
            for (size_t x = 0; x < ca.size(); ++x)
                for (size_t y = 0; y < cb.size(); ++y) {
                    double *e = &pp.E[(x * cb.size() + y) * (size_t) (D * D * D)];
                    for (int t = 0; t <= ca[x].x + cb[y].x; ++t)
                        for (int u = 0; u <= ca[x].y + cb[y].y; ++u)
                            for (int v = 0; v <= ca[x].z + cb[y].z; ++v)
                                e[(t * D + u) * D + v] = E[0](ca[x].x, cb[y].x, t) * E[1](ca[x].y, cb[y].y, u) * E[2](ca[x].z, cb[y].z, v);
                }
            sp.prims.push_back(std::move(pp));
        }
    return sp;
}

}// namespace

double contracted_norm(const Shell &s) {
    std::vector<double> S;
    overlap_cart(s, s, S);
    return S[0];// <x^l|x^l> of the contracted function built from normalised primitives
}

Matrix compute_1body_ints(const BasisSet &obs, Operator1 op, const std::vector<Atom> &atoms) {
    const auto &sh = obs.shells();
    const auto s2bf = obs.shell2bf();
    const size_t n = obs.nbf();
    Matrix result((long) n, (long) n);
    std::vector<std::vector<std::vector<double>>> Ts;
    for (const Shell &s : sh) Ts.push_back(shell_transform(s));
#pragma omp parallel for schedule(dynamic)
    for (long s1 = 0; s1 < (long) sh.size(); ++s1)
        for (long s2 = 0; s2 <= s1; ++s2) {
            const Shell &A = sh[(size_t) s1], &B = sh[(size_t) s2];
            const auto ca = cart_components(A.l), cb = cart_components(B.l);
            std::vector<double> blk(ca.size() * cb.size(), 0.0), pure;
            if (op == Operator1::overlap) {
                overlap_cart(A, B, blk);
            } else {
                for (size_t i = 0; i < A.alpha.size(); ++i)
                    for (size_t j = 0; j < B.alpha.size(); ++j) {
                        const double a = A.alpha[i], b = B.alpha[j], p = a + b;
                        const double w = A.coef[i] * prim_norm(a, A.l) * B.coef[j] * prim_norm(b, B.l);
                        ETable E[3];
                        for (int d = 0; d < 3; ++d) E[d] = hermite_E(A.l, B.l + 2, a, b, A.O[(size_t) d] - B.O[(size_t) d]);
                        if (op == Operator1::kinetic) {
                            const double f = std::pow(kPi / p, 1.5);
                            auto d2 = [&](int d, int ii, int jj) {// <i| d^2/dx^2 |j> / sqrt(pi/p)
                                return (jj > 1 ? jj * (jj - 1) * E[d](ii, jj - 2, 0) : 0.0) - 2.0 * b * (2 * jj + 1) * E[d](ii, jj, 0) +
                                       4.0 * b * b * E[d](ii, jj + 2, 0);
                            };
                            for (size_t x = 0; x < ca.size(); ++x)
                                for (size_t y = 0; y < cb.size(); ++y) {
                                    const double sx = E[0](ca[x].x, cb[y].x, 0), sy = E[1](ca[x].y, cb[y].y, 0), sz = E[2](ca[x].z, cb[y].z, 0);
                                    blk[x * cb.size() + y] += -0.5 * w * f *
                                                              (d2(0, ca[x].x, cb[y].x) * sy * sz + sx * d2(1, ca[x].y, cb[y].y) * sz +
                                                               sx * sy * d2(2, ca[x].z, cb[y].z));
                                }
                        } else {// nuclear attraction
                            double P[3];
                            for (int d = 0; d < 3; ++d) P[d] = (a * A.O[(size_t) d] + b * B.O[(size_t) d]) / p;
                            const int L = A.l + B.l;
                            RTable R;
                            std::vector<double> work;
                            for (const Atom &at : atoms) {
                                const double PC[3] = {P[0] - at.x, P[1] - at.y, P[2] - at.z};
                                hermite_R(L, p, PC, R, work);
                                const double f = -at.atomic_number * 2.0 * kPi / p * w;
                                for (size_t x = 0; x < ca.size(); ++x)
                                    for (size_t y = 0; y < cb.size(); ++y) {
                                        double acc = 0.0;
                                        for (int t = 0; t <= ca[x].x + cb[y].x; ++t)
                                            for (int u = 0; u <= ca[x].y + cb[y].y; ++u)
                                                for (int v = 0; v <= ca[x].z + cb[y].z; ++v)
                                                    acc += E[0](ca[x].x, cb[y].x, t) * E[1](ca[x].y, cb[y].y, u) * E[2](ca[x].z, cb[y].z, v) * R(t, u, v);
                                        blk[x * cb.size() + y] += f * acc;
                                    }
                            }
                        }
                    }
            }
            transform2(Ts[(size_t) s1], Ts[(size_t) s2], blk, ca.size(), cb.size(), pure);
            const size_t n1 = A.size(), n2 = B.size();
            for (size_t x = 0; x < n1; ++x)
                for (size_t y = 0; y < n2; ++y) {
                    result((long) (s2bf[(size_t) s1] + x), (long) (s2bf[(size_t) s2] + y)) = pure[x * n2 + y];
                    result((long) (s2bf[(size_t) s2] + y), (long) (s2bf[(size_t) s1] + x)) = pure[x * n2 + y];
                }
        }
    return result;
}

double compute_enuc(const std::vector<Atom> &atoms) {// hf.cpp:21-34
    double e = 0.0;
    for (size_t i = 0; i < atoms.size(); ++i)
        for (size_t j = i + 1; j < atoms.size(); ++j) {
            const double dx = atoms[i].x - atoms[j].x, dy = atoms[i].y - atoms[j].y, dz = atoms[i].z - atoms[j].z;
            e += atoms[i].atomic_number * atoms[j].atomic_number / std::sqrt(dx * dx + dy * dy + dz * dz);
        }
    return e;
}

// integrals.cpp:12-89: full (pq|rs) tensor from the permutationally unique shell quartets
DTensor eri_ao_tensor(const BasisSet &obs) {
    const auto &sh = obs.shells();
    const auto s2bf = obs.shell2bf();
    const long n = (long) obs.nbf(), ns = (long) sh.size();
    DTensor ao(n, n, n, n);
    std::vector<std::vector<std::vector<double>>> Ts;
    for (const Shell &s : sh) Ts.push_back(shell_transform(s));
    // pair data for s1 >= s2
    std::vector<ShellPair> pairs((size_t) (ns * ns));
#pragma omp parallel for schedule(dynamic)
    for (long s1 = 0; s1 < ns; ++s1)
        for (long s2 = 0; s2 <= s1; ++s2) pairs[(size_t) (s1 * ns + s2)] = make_pair_data(sh[(size_t) s1], sh[(size_t) s2]);
    std::vector<std::pair<long, long>> plist;
    for (long s1 = 0; s1 < ns; ++s1)
        for (long s2 = 0; s2 <= s1; ++s2) plist.push_back({s1, s2});
    const double pref0 = 2.0 * std::pow(kPi, 2.5);
#pragma omp parallel
    {
        RTable R;
        std::vector<double> work, G, cartblk, tmp1, tmp2;
#pragma omp for schedule(dynamic)
        for (long pi = 0; pi < (long) plist.size(); ++pi)
            for (long qi = 0; qi <= pi; ++qi) {
                const long s1 = plist[(size_t) pi].first, s2 = plist[(size_t) pi].second, s3 = plist[(size_t) qi].first,
                           s4 = plist[(size_t) qi].second;
                const Shell &A = sh[(size_t) s1], &B = sh[(size_t) s2], &C = sh[(size_t) s3], &Dd = sh[(size_t) s4];
                const size_t nca = (size_t) A.ncart(), ncb = (size_t) B.ncart(), ncc = (size_t) C.ncart(), ncd = (size_t) Dd.ncart();
                const size_t nab = nca * ncb, ncdp = ncc * ncd;
                const int Lab = A.l + B.l, Lcd = C.l + Dd.l, L = Lab + Lcd;
                const int Db = Lab + 1, Dk = Lcd + 1, Dr = L + 1;
                cartblk.assign(nab * ncdp, 0.0);
                const auto ck = cart_components(C.l), dk = cart_components(Dd.l);
                for (const PrimPair &bra : pairs[(size_t) (s1 * ns + s2)].prims)
                    for (const PrimPair &ket : pairs[(size_t) (s3 * ns + s4)].prims) {
                        const double p = bra.p, q = ket.p, alpha = p * q / (p + q);
                        const double PQ[3] = {bra.P[0] - ket.P[0], bra.P[1] - ket.P[1], bra.P[2] - ket.P[2]};
                        hermite_R(L, alpha, PQ, R, work);
                        const double pref = pref0 / (p * q * std::sqrt(p + q)) * bra.w * ket.w;
                        G.assign((size_t) (Db * Db * Db), 0.0);
                        for (size_t kc = 0; kc < ncdp; ++kc) {
                            const double *ek = &ket.E[kc * (size_t) (Dk * Dk * Dk)];
                            const size_t ic = kc / ncd, id = kc % ncd;
                            const int tx = ck[ic].x + dk[id].x, ty = ck[ic].y + dk[id].y, tz = ck[ic].z + dk[id].z;
                            for (int t = 0; t <= Lab; ++t)
                                for (int u = 0; u <= Lab - t; ++u)
                                    for (int v = 0; v <= Lab - t - u; ++v) {
                                        double acc = 0.0;
                                        for (int a = 0; a <= tx; ++a)
                                            for (int b = 0; b <= ty; ++b)
                                                for (int c = 0; c <= tz; ++c) {
                                                    const double e = ek[(a * Dk + b) * Dk + c];
                                                    acc += (((a + b + c) & 1) ? -e : e) * R.r[(size_t) (((t + a) * Dr + (u + b)) * Dr + (v + c))];
                                                }
                                        G[(size_t) ((t * Db + u) * Db + v)] = acc;
                                    }
                            for (size_t kb = 0; kb < nab; ++kb) {
                                const double *eb = &bra.E[kb * (size_t) (Db * Db * Db)];
                                double acc = 0.0;
                                for (int t = 0; t <= Lab; ++t)
                                    for (int u = 0; u <= Lab - t; ++u)
                                        for (int v = 0; v <= Lab - t - u; ++v)
                                            acc += eb[(t * Db + u) * Db + v] * G[(size_t) ((t * Db + u) * Db + v)];
                                cartblk[kb * ncdp + kc] += pref * acc;
                            }
                        }
                    }
                // Cartesian -> shell functions, index by index
                const auto &Ta = Ts[(size_t) s1], &Tb = Ts[(size_t) s2], &Tc = Ts[(size_t) s3], &Td = Ts[(size_t) s4];
                const size_t n1 = Ta.size(), n2 = Tb.size(), n3 = Tc.size(), n4 = Td.size();
                // (ca cb | cc cd) -> (ca cb | p3 p4)
                tmp1.assign(nab * n3 * n4, 0.0);
                for (size_t kb = 0; kb < nab; ++kb)
                    for (size_t p3 = 0; p3 < n3; ++p3)
                        for (size_t p4 = 0; p4 < n4; ++p4) {
                            double acc = 0.0;
                            for (size_t x = 0; x < ncc; ++x) {
                                if (Tc[p3][x] == 0.0) continue;
                                for (size_t y = 0; y < ncd; ++y) acc += Tc[p3][x] * Td[p4][y] * cartblk[kb * ncdp + x * ncd + y];
                            }
                            tmp1[(kb * n3 + p3) * n4 + p4] = acc;
                        }
                tmp2.assign(n1 * n2 * n3 * n4, 0.0);
                for (size_t p1 = 0; p1 < n1; ++p1)
                    for (size_t p2 = 0; p2 < n2; ++p2)
                        for (size_t x = 0; x < nca; ++x) {
                            if (Ta[p1][x] == 0.0) continue;
                            for (size_t y = 0; y < ncb; ++y) {
                                const double f = Ta[p1][x] * Tb[p2][y];
                                if (f == 0.0) continue;
                                const double *src = &tmp1[(x * ncb + y) * n3 * n4];
                                double *dst = &tmp2[(p1 * n2 + p2) * n3 * n4];
                                for (size_t k = 0; k < n3 * n4; ++k) dst[k] += f * src[k];
                            }
                        }
                for (size_t f1 = 0; f1 < n1; ++f1)
                    for (size_t f2 = 0; f2 < n2; ++f2)
                        for (size_t f3 = 0; f3 < n3; ++f3)
                            for (size_t f4 = 0; f4 < n4; ++f4) {
                                const double val = tmp2[((f1 * n2 + f2) * n3 + f3) * n4 + f4];
                                const long b1 = (long) (s2bf[(size_t) s1] + f1), b2 = (long) (s2bf[(size_t) s2] + f2),
                                           b3 = (long) (s2bf[(size_t) s3] + f3), b4 = (long) (s2bf[(size_t) s4] + f4);
                                ao(b1, b2, b3, b4) = val, ao(b1, b2, b4, b3) = val, ao(b2, b1, b3, b4) = val, ao(b2, b1, b4, b3) = val;
                                ao(b3, b4, b1, b2) = val, ao(b3, b4, b2, b1) = val, ao(b4, b3, b1, b2) = val, ao(b4, b3, b2, b1) = val;
                            }
            }
    }
    return ao;
}
