// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline).
//
// CPU restatement of the post-HF hot path of ajay-mk/QC-P2:
//   src/integrals.cpp:95-207, src/mp2.cpp:11-42, src/cc.cpp:8-560.
// Nothing in the product path (qc-p2_b200/) may include, link or call this file;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs use it.
//
// PARITY PIN: the reference's own sources (src/{integrals,mp2,cc,hf,general}.cpp, unmodified,
// compiled where they lie by oracle/ref_shim/Makefile against API stand-ins for the absent
// third-party headers, output oracle/_ref/) are run beside this restatement by
// tests/test_ref_pins_oracle.py: every function of SURVEY.md section 8a, RHF and as-written
// UHF, agrees to rounding (bit-identical energies on the seeded cases), and the reference
// program's results on BASELINE.json configs 1, 2, 4 are committed as
// tests/golden/reference_energies.json.  What remains restated rather than run is the
// third-party arithmetic itself: BTAS (github.com/ValeevGroup/btas, unpinned master,
// CMakeLists.txt:9-13) is not in this image, so `contract(alpha,A,ia,B,ib,beta,C,ic)` means,
// here and in the shim, its published C[ic] = beta*C[ic] + alpha * sum_{shared} A[ia]*B[ib].
// External anchor: the Crawford ProgrammingProjects water/STO-3G known answers (cc.cpp:453),
// tests/test_known_answers.py.
//
// Written as an index-string (einsum-like) engine so each reference contraction is one
// line whose labels can be compared with the cited reference line.

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// Dense row-major tensor (last index fastest), zero-initialised on construction.
// Zero-init matters: the reference accumulates into freshly constructed tensors
// (e.g. integrals.cpp:109, cc.cpp:192, 300-312, 472-500).
struct Tensor {
    std::vector<long> ext;
    std::vector<double> v;
    Tensor() {}
    explicit Tensor(const std::vector<long> &e) : ext(e) {
        long n = 1;
        for (long x : e) n *= x;
        v.assign((size_t) n, 0.0);
    }
    Tensor(const std::vector<long> &e, const double *src) : Tensor(e) {
        if (src) std::memcpy(v.data(), src, v.size() * sizeof(double));
    }
    int rank() const { return (int) ext.size(); }
    long extent(int k) const { return ext[(size_t) k]; }
    long size() const { return (long) v.size(); }
    double *data() { return v.data(); }
    const double *data() const { return v.data(); }
    template<class... I>
    long off(I... idx) const {
        const long ix[] = {(long) idx...};
        long o = 0;
        for (size_t k = 0; k < sizeof...(I); ++k) o = o * ext[k] + ix[k];
        return o;
    }
    template<class... I>
    double &operator()(I... idx) { return v[(size_t) off(idx...)]; }
    template<class... I>
    double operator()(I... idx) const { return v[(size_t) off(idx...)]; }
    Tensor &operator+=(const Tensor &o) {
        assert(o.v.size() == v.size());
        for (size_t k = 0; k < v.size(); ++k) v[k] += o.v[k];
        return *this;
    }
    Tensor &operator-=(const Tensor &o) {
        assert(o.v.size() == v.size());
        for (size_t k = 0; k < v.size(); ++k) v[k] -= o.v[k];
        return *this;
    }
};

static Tensor T2d(long a, long b) { return Tensor(std::vector<long>{a, b}); }
static Tensor T4d(long a, long b, long c, long d) { return Tensor(std::vector<long>{a, b, c, d}); }

// ---------------------------------------------------------------------------------
// contract: C[ic] = beta*C[ic] + alpha * sum_K A[ia]*B[ib]      (BTAS semantics)
// Implemented as gather-to-matrix, blocked OpenMP GEMM, scatter-accumulate.
// ---------------------------------------------------------------------------------
static void gemm_acc(long M, long N, long K, const double *A, const double *B, double *C) {
    // C[M,N] = A[M,K] * B[K,N], all row-major
    const long JB = 1024;
    const long njb = (N + JB - 1) / JB;
#pragma omp parallel for collapse(2) schedule(static)
    for (long i = 0; i < M; ++i) {
        for (long jb = 0; jb < njb; ++jb) {
            const long j0 = jb * JB, j1 = std::min(N, j0 + JB);
            double *c = C + i * N;
            for (long j = j0; j < j1; ++j) c[j] = 0.0;
            for (long k = 0; k < K; ++k) {
                const double a = A[i * K + k];
                const double *b = B + k * N;
                for (long j = j0; j < j1; ++j) c[j] += a * b[j];
            }
        }
    }
}

struct Group {
    std::vector<int> posA, posB, posC;// positions of the group's labels in each operand (-1 if absent)
    std::vector<long> ext;
    long total = 1;
};

static void contract(double alpha, const Tensor &A, const char *ia, const Tensor &B, const char *ib, double beta,
                     Tensor &C, const char *ic) {
    const std::string sa(ia), sb(ib), sc(ic);
    if ((int) sa.size() != A.rank() || (int) sb.size() != B.rank() || (int) sc.size() != C.rank())
        throw std::invalid_argument("contract: label/rank mismatch " + sa + "," + sb + "->" + sc);
    Group gm, gn, gk;
    for (size_t p = 0; p < sc.size(); ++p) {
        const char l = sc[p];
        const size_t pa = sa.find(l), pb = sb.find(l);
        if (pa != std::string::npos && pb == std::string::npos) {
            gm.posA.push_back((int) pa), gm.posC.push_back((int) p), gm.ext.push_back(C.extent((int) p));
            if (A.extent((int) pa) != C.extent((int) p)) throw std::invalid_argument("contract: extent mismatch (M)");
        } else if (pb != std::string::npos && pa == std::string::npos) {
            gn.posB.push_back((int) pb), gn.posC.push_back((int) p), gn.ext.push_back(C.extent((int) p));
            if (B.extent((int) pb) != C.extent((int) p)) throw std::invalid_argument("contract: extent mismatch (N)");
        } else
            throw std::invalid_argument("contract: unsupported label pattern " + sa + "," + sb + "->" + sc);
    }
    for (size_t p = 0; p < sa.size(); ++p) {
        const char l = sa[p];
        if (sc.find(l) != std::string::npos) continue;
        const size_t pb = sb.find(l);
        if (pb == std::string::npos) throw std::invalid_argument("contract: dangling label in A");
        if (A.extent((int) p) != B.extent((int) pb)) throw std::invalid_argument("contract: extent mismatch (K)");
        gk.posA.push_back((int) p), gk.posB.push_back((int) pb), gk.ext.push_back(A.extent((int) p));
    }
    for (long e : gm.ext) gm.total *= e;
    for (long e : gn.ext) gn.total *= e;
    for (long e : gk.ext) gk.total *= e;
    const long M = gm.total, N = gn.total, K = gk.total;

    auto strides = [](const Tensor &t) {
        std::vector<long> s((size_t) t.rank(), 1);
        for (int k = t.rank() - 2; k >= 0; --k) s[(size_t) k] = s[(size_t) k + 1] * t.extent(k + 1);
        return s;
    };
    const auto stA = strides(A), stB = strides(B), stC = strides(C);
    // offset tables: for every composite index of a group, the offset it contributes inside an operand
    auto table = [](const std::vector<long> &ext, const std::vector<int> &pos, const std::vector<long> &st) {
        long tot = 1;
        for (long e : ext) tot *= e;
        std::vector<long> t((size_t) tot, 0);
        for (long x = 0; x < tot; ++x) {
            long r = x, o = 0;
            for (int d = (int) ext.size() - 1; d >= 0; --d) {
                o += (r % ext[(size_t) d]) * st[(size_t) pos[(size_t) d]];
                r /= ext[(size_t) d];
            }
            t[(size_t) x] = o;
        }
        return t;
    };
    const auto mA = table(gm.ext, gm.posA, stA), mC = table(gm.ext, gm.posC, stC);
    const auto nB = table(gn.ext, gn.posB, stB), nC = table(gn.ext, gn.posC, stC);
    const auto kA = table(gk.ext, gk.posA, stA), kB = table(gk.ext, gk.posB, stB);

    std::vector<double> Am((size_t) (M * K)), Bm((size_t) (K * N)), Cm((size_t) (M * N));
#pragma omp parallel for schedule(static)
    for (long m = 0; m < M; ++m)
        for (long k = 0; k < K; ++k) Am[(size_t) (m * K + k)] = A.v[(size_t) (mA[(size_t) m] + kA[(size_t) k])];
#pragma omp parallel for schedule(static)
    for (long k = 0; k < K; ++k)
        for (long n = 0; n < N; ++n) Bm[(size_t) (k * N + n)] = B.v[(size_t) (kB[(size_t) k] + nB[(size_t) n])];
    gemm_acc(M, N, K, Am.data(), Bm.data(), Cm.data());
#pragma omp parallel for schedule(static)
    for (long m = 0; m < M; ++m)
        for (long n = 0; n < N; ++n) {
            double &c = C.v[(size_t) (mC[(size_t) m] + nC[(size_t) n])];
            c = beta * c + alpha * Cm[(size_t) (m * N + n)];
        }
}

// ---------------------------------------------------------------------------------
// integrals.cpp
// ---------------------------------------------------------------------------------

// integrals.cpp:95-134.  Indices k,l are rotated by Coeff1, indices i,j by Coeff2
// (integrals.cpp:121-129); coefficient matrices are row-major [AO][MO].
static Tensor transform_ao_mo(const Tensor &pq_rs, const Tensor &C1, const Tensor &C2) {
    const long n = pq_rs.extent(0);
    Tensor pq_rl = T4d(n, n, n, n), pq_kl = T4d(n, n, n, n), pj_kl = T4d(n, n, n, n), ij_kl = T4d(n, n, n, n);
    contract(1.0, pq_rs, "pqrs", C1, "sl", 1.0, pq_rl, "pqrl");// :121
    contract(1.0, pq_rl, "pqrl", C1, "rk", 1.0, pq_kl, "pqkl");// :123
    contract(1.0, C2, "qj", pq_kl, "pqkl", 1.0, pj_kl, "pjkl");// :126
    contract(1.0, C2, "pi", pj_kl, "pjkl", 1.0, ij_kl, "ijkl");// :129
    return ij_kl;
}

// integrals.cpp:136-167.  Even spin-orbital = alpha, odd = beta; <ij||kl> = (ik|jl) - (jk|il).
// The 10 spin patterns not listed stay at their zero initial value (Q2); the alpha-beta
// branches read mo_ints_ab exactly as written, including the UHF label swap (Q5).
static Tensor transform_to_so(const Tensor &aa, const Tensor &bb, const Tensor &ab) {
    const long n = aa.extent(0) * 2;
    Tensor so = T4d(n, n, n, n);
#pragma omp parallel for collapse(2) schedule(static)
    for (long i = 0; i < n; ++i)
        for (long j = 0; j < n; ++j)
            for (long k = 0; k < n; ++k)
                for (long l = 0; l < n; ++l) {
                    const int si = (int) (i & 1), sj = (int) (j & 1), sk = (int) (k & 1), sl = (int) (l & 1);
                    const long I = i / 2, J = j / 2, Kk = k / 2, L = l / 2;
                    double val = 0.0;
                    bool hit = true;
                    if (!si && !sk && !sj && !sl) val = aa(I, Kk, J, L) - aa(J, Kk, I, L);     // :143-145
                    else if (si && sk && sj && sl) val = bb(I, Kk, J, L) - bb(J, Kk, I, L);    // :147-149
                    else if (!si && !sk && sj && sl) val = ab(I, Kk, J, L);                   // :151-152
                    else if (si && sk && !sj && !sl) val = ab(I, Kk, J, L);                   // :154-155
                    else if (si && !sk && !sj && sl) val = -ab(J, Kk, I, L);                  // :157-158
                    else if (!si && sk && sj && !sl) val = -ab(J, Kk, I, L);                  // :160-161
                    else hit = false;
                    if (hit) so(i, j, k, l) = val;
                }
    return so;
}

// integrals.cpp:169-181
static Tensor make_denom(const Tensor &F_spin, long no, long nv) {
    Tensor d = T4d(no, no, nv, nv);
    for (long i = 0; i < no; ++i)
        for (long j = 0; j < no; ++j)
            for (long a = 0; a < nv; ++a)
                for (long b = 0; b < nv; ++b)
                    d(i, j, a, b) = F_spin(i, i) + F_spin(j, j) - F_spin(no + a, no + a) - F_spin(no + b, no + b);
    return d;
}

// integrals.cpp:184-207.  spec is a 4-letter string over {o,v}.
static Tensor slice_ints(const Tensor &so, long no, long nv, const char *spec) {
    long n[4], m[4];
    for (int d = 0; d < 4; ++d) {
        n[d] = (spec[d] == 'o') ? no : (spec[d] == 'v' ? nv : 0);
        m[d] = (spec[d] == 'v') ? no : 0;
    }
    Tensor s = T4d(n[0], n[1], n[2], n[3]);
#pragma omp parallel for schedule(static)
    for (long p = 0; p < n[0]; ++p)
        for (long q = 0; q < n[1]; ++q)
            for (long r = 0; r < n[2]; ++r)
                for (long t = 0; t < n[3]; ++t) s(p, q, r, t) = so(p + m[0], q + m[1], r + m[2], t + m[3]);
    return s;
}

// ---------------------------------------------------------------------------------
// mp2.cpp
// ---------------------------------------------------------------------------------

// mp2.cpp:11-21: diagonal spin-orbital Fock matrix, alpha on even, beta on odd positions.
static Tensor make_so_moes(const double *eps_a, const double *eps_b, long nao) {
    const long n = 2 * nao;
    Tensor e = T2d(n, n);
    for (long i = 0; i < n; ++i) e(i, i) = (i % 2 == 0) ? eps_a[i / 2] : eps_b[i / 2];
    return e;
}

// mp2.cpp:23-42: T = <ij||ab>/D and E = 1/4 sum <ij||ab>^2 / D, accumulated in loop order.
static double run_mp2(const Tensor &oovv, const Tensor &denom, Tensor &T) {
    const long no = oovv.extent(0), nv = oovv.extent(2);
    T = T4d(no, no, nv, nv);
    double e = 0.0;
    for (long x = 0; x < oovv.size(); ++x) {
        T.v[(size_t) x] = oovv.v[(size_t) x] / denom.v[(size_t) x];
        e += 0.25 * (oovv.v[(size_t) x] * oovv.v[(size_t) x]) / denom.v[(size_t) x];
    }
    return e;
}

// ---------------------------------------------------------------------------------
// cc.cpp
// ---------------------------------------------------------------------------------

struct Ints {// integrals.h:11-23, cc.cpp:8-23
    Tensor oooo, ovoo, oovo, ooov, ovov, ovvo, oovv, vovv, ovvv, vvvo, vvvv;
};
static const char *kIntNames[11] = {"oooo", "ovoo", "oovo", "ooov", "ovov", "ovvo",
                                    "oovv", "vovv", "ovvv", "vvvo", "vvvv"};
static Tensor *int_field(Ints &s, int k) {
    Tensor *f[11] = {&s.oooo, &s.ovoo, &s.oovo, &s.ooov, &s.ovov, &s.ovvo, &s.oovv, &s.vovv, &s.ovvv, &s.vvvo, &s.vvvv};
    return f[k];
}
static Ints get_integrals(const Tensor &so, long no, long nv) {
    Ints s;
    for (int k = 0; k < 11; ++k) *int_field(s, k) = slice_ints(so, no, nv, kIntNames[k]);
    return s;
}

// cc.cpp:25-34 (even-n1 branch; the odd branch is unreachable, SURVEY Q6, but kept literal)
static Tensor make_fock(const double *e1, const double *e2, long n1, long n2) {
    Tensor f = T2d(n2 - n1, n2 - n1);
    for (long i = 0; i < n2 - n1; ++i) {
        const long base = (n1 % 2 == 0) ? n1 / 2 + i / 2 : n1 / 2 + i / 2 + 1;
        f(i, i) = (i % 2 == 0) ? e1[base] : e2[base];
    }
    return f;
}

struct Moes {// cc.h:17-19
    Tensor F_ii, F_ia, F_aa;
};

// cc.cpp:73-100
static Tensor make_D_ia(const Moes &m) {
    const long no = m.F_ii.extent(0), nv = m.F_aa.extent(0);
    Tensor d = T2d(no, nv);
    for (long i = 0; i < no; ++i)
        for (long a = 0; a < nv; ++a) d(i, a) = m.F_ii(i, i) - m.F_aa(a, a);
    return d;
}
static Tensor make_D_ijab(const Moes &m) {
    const long no = m.F_ii.extent(0), nv = m.F_aa.extent(0);
    Tensor d = T4d(no, no, nv, nv);
    for (long i = 0; i < no; ++i)
        for (long j = 0; j < no; ++j)
            for (long a = 0; a < nv; ++a)
                for (long b = 0; b < nv; ++b) d(i, j, a, b) = m.F_ii(i, i) + m.F_ii(j, j) - m.F_aa(a, a) - m.F_aa(b, b);
    return d;
}

// cc.cpp:125-172
static Tensor make_tau_like(const Tensor &Ts, const Tensor &Td, double w) {
    const long no = Ts.extent(0), nv = Ts.extent(1);
    Tensor t = T4d(no, no, nv, nv);
    for (long i = 0; i < no; ++i)
        for (long j = 0; j < no; ++j)
            for (long a = 0; a < nv; ++a)
                for (long b = 0; b < nv; ++b)
                    t(i, j, a, b) = Td(i, j, a, b) + w * (Ts(i, a) * Ts(j, b) - Ts(i, b) * Ts(j, a));
    return t;
}
static Tensor make_tau(const Tensor &Ts, const Tensor &Td) {// Stanton eq. 9, cc.cpp:125-139
    const long no = Ts.extent(0), nv = Ts.extent(1);
    Tensor t = T4d(no, no, nv, nv);
    for (long i = 0; i < no; ++i)
        for (long j = 0; j < no; ++j)
            for (long a = 0; a < nv; ++a)
                for (long b = 0; b < nv; ++b) t(i, j, a, b) = Td(i, j, a, b) + Ts(i, a) * Ts(j, b) - Ts(i, b) * Ts(j, a);
    return t;
}
static Tensor make_tau_bar(const Tensor &Ts, const Tensor &Td) { return make_tau_like(Ts, Td, 0.5); }// eq. 10, :141-156
static Tensor multiply_Ts(const Tensor &Ts) {// cc.cpp:158-172, layout [j,f,n,b]
    const long no = Ts.extent(0), nv = Ts.extent(1);
    Tensor r = T4d(no, nv, no, nv);
    for (long j = 0; j < no; ++j)
        for (long f = 0; f < nv; ++f)
            for (long n = 0; n < no; ++n)
                for (long b = 0; b < nv; ++b) r(j, f, n, b) = Ts(j, f) * Ts(n, b);
    return r;
}

struct Inter {// cc.h:14-16
    Tensor F_ae, F_mi, F_me, W_mnij, W_mbej, W_abef;
};

// cc.cpp:175-255 (Stanton eqs. 3-8).  Label strings are those of the cited lines.
static Inter update_intermediates(const Tensor &Ts, const Tensor &Td, const Ints &I, const Moes &mo) {
    const long no = Ts.extent(0), nv = Ts.extent(1);
    Inter X;
    X.F_ae = T2d(nv, nv);
    contract(-0.5, mo.F_ia, "me", Ts, "ma", 1.0, X.F_ae, "ae");                        // :194
    contract(1.0, Ts, "mf", I.ovvv, "mafe", 1.0, X.F_ae, "ae");                        // :195
    contract(-0.5, make_tau_bar(Ts, Td), "mnaf", I.oovv, "mnef", 1.0, X.F_ae, "ae");   // :196
    for (long a = 0; a < nv; ++a)
        for (long e = 0; e < nv; ++e) X.F_ae(a, e) += (1 - (a == e)) * mo.F_aa(a, e);  // :198-202

    X.F_mi = T2d(no, no);
    contract(0.5, Ts, "ie", mo.F_ia, "me", 1.0, X.F_mi, "mi");                         // :207
    contract(1.0, Ts, "ne", I.ooov, "mnie", 1.0, X.F_mi, "mi");                        // :208
    contract(0.5, make_tau_bar(Ts, Td), "inef", I.oovv, "mnef", 1.0, X.F_mi, "mi");    // :209
    for (long m = 0; m < no; ++m)
        for (long i = 0; i < no; ++i) X.F_mi(m, i) += (1 - (m == i)) * mo.F_ii(m, i);  // :212-216

    X.F_me = T2d(no, nv);
    contract(1.0, Ts, "nf", I.oovv, "mnef", 1.0, X.F_me, "me");                        // :220
    X.F_me += mo.F_ia;                                                                 // :221-225

    X.W_mnij = T4d(no, no, no, no);
    contract(1.0, Ts, "je", I.ooov, "mnie", 1.0, X.W_mnij, "mnij");                    // :229
    contract(-1.0, Ts, "ie", I.ooov, "mnje", 1.0, X.W_mnij, "mnij");                   // :230
    contract(0.25, make_tau(Ts, Td), "ijef", I.oovv, "mnef", 1.0, X.W_mnij, "mnij");   // :231
    X.W_mnij += I.oooo;                                                                // :232

    X.W_abef = T4d(nv, nv, nv, nv);
    contract(-1.0, Ts, "mb", I.vovv, "amef", 1.0, X.W_abef, "abef");                   // :237
    contract(1.0, Ts, "ma", I.vovv, "bmef", 1.0, X.W_abef, "abef");                    // :238
    contract(0.25, make_tau(Ts, Td), "mnab", I.oovv, "mnef", 1.0, X.W_abef, "abef");   // :239
    X.W_abef += I.vvvv;                                                                // :240

    X.W_mbej = T4d(no, nv, nv, no);
    contract(1.0, Ts, "jf", I.ovvv, "mbef", 1.0, X.W_mbej, "mbej");                    // :245
    contract(-1.0, Ts, "nb", I.oovo, "mnej", 1.0, X.W_mbej, "mbej");                   // :246
    contract(-0.5, Td, "jnfb", I.oovv, "mnef", 1.0, X.W_mbej, "mbej");                 // :247
    contract(-1.0, multiply_Ts(Ts), "jfnb", I.oovv, "mnef", 1.0, X.W_mbej, "mbej");    // :249
    X.W_mbej += I.ovvo;                                                                // :250
    return X;
}

// cc.cpp:257-292 (Stanton eq. 1)
static Tensor make_T1(const Tensor &Ts, const Tensor &Td, const Ints &I, const Inter &X, const Tensor &D_ia,
                      const Moes &mo) {
    const long no = Ts.extent(0), nv = Ts.extent(1);
    Tensor r = T2d(no, nv);
    contract(1.0, Ts, "ie", X.F_ae, "ae", 1.0, r, "ia");       // :272
    contract(-1.0, Ts, "ma", X.F_mi, "mi", 1.0, r, "ia");      // :273
    contract(1.0, Td, "imae", X.F_me, "me", 1.0, r, "ia");     // :274
    contract(-1.0, Ts, "nf", I.ovov, "naif", 1.0, r, "ia");    // :275
    contract(-0.5, Td, "imef", I.ovvv, "maef", 1.0, r, "ia");  // :276
    contract(-0.5, Td, "mnae", I.oovo, "nmei", 1.0, r, "ia");  // :277
    r += mo.F_ia;                                              // :280
    Tensor t = T2d(no, nv);
    for (long x = 0; x < r.size(); ++x) t.v[(size_t) x] = r.v[(size_t) x] / D_ia.v[(size_t) x];// :283-288
    return t;
}

// cc.cpp:294-387 (Stanton eq. 2)
static Tensor make_T2(const Tensor &Ts, const Tensor &Td, const Ints &I, const Inter &X, const Tensor &D_ijab) {
    const long no = Ts.extent(0), nv = Ts.extent(1);
    Tensor r = T4d(no, no, nv, nv);
    r += I.oovv;                                                       // :312
    contract(1.0, Td, "ijae", X.F_ae, "be", 1.0, r, "ijab");           // :316
    contract(-1.0, Td, "ijbe", X.F_ae, "ae", 1.0, r, "ijab");          // :317
    Tensor d1 = T2d(nv, nv);
    contract(1.0, Ts, "mb", X.F_me, "me", 1.0, d1, "be");              // :322
    contract(-0.5, Td, "ijae", d1, "be", 1.0, r, "ijab");              // :324
    d1 -= d1;                                                          // :326
    contract(1.0, Ts, "ma", X.F_me, "me", 1.0, d1, "ae");              // :327
    contract(0.5, Td, "ijbe", d1, "ae", 1.0, r, "ijab");               // :328
    contract(-1.0, Td, "imab", X.F_mi, "mj", 1.0, r, "ijab");          // :333
    contract(1.0, Td, "jmab", X.F_mi, "mi", 1.0, r, "ijab");           // :334
    Tensor d2 = T2d(no, no);
    contract(1.0, Ts, "je", X.F_me, "me", 1.0, d2, "jm");              // :339
    contract(-0.5, Td, "imab", d2, "jm", 1.0, r, "ijab");              // :340
    d2 -= d2;                                                          // :341
    contract(1.0, Ts, "ie", X.F_me, "me", 1.0, d2, "im");              // :342
    contract(0.5, Td, "jmab", d2, "im", 1.0, r, "ijab");               // :343
    contract(0.5, make_tau(Ts, Td), "mnab", X.W_mnij, "mnij", 1.0, r, "ijab");// :346
    contract(0.5, make_tau(Ts, Td), "ijef", X.W_abef, "abef", 1.0, r, "ijab");// :348
    contract(1.0, Td, "imae", X.W_mbej, "mbej", 1.0, r, "ijab");       // :354
    contract(-1.0, Td, "jmae", X.W_mbej, "mbei", 1.0, r, "ijab");      // :355
    contract(-1.0, Td, "imbe", X.W_mbej, "maej", 1.0, r, "ijab");      // :356
    contract(1.0, Td, "jmbe", X.W_mbej, "maei", 1.0, r, "ijab");       // :357
    contract(1.0, multiply_Ts(Ts), "iema", I.ovov, "mbje", 1.0, r, "ijab"); // :360
    contract(-1.0, multiply_Ts(Ts), "jema", I.ovov, "mbie", 1.0, r, "ijab");// :361
    contract(-1.0, multiply_Ts(Ts), "iemb", I.ovov, "maje", 1.0, r, "ijab");// :362
    contract(1.0, multiply_Ts(Ts), "jemb", I.ovov, "maie", 1.0, r, "ijab"); // :363
    contract(1.0, Ts, "ie", I.vvvo, "abej", 1.0, r, "ijab");           // :366
    contract(-1.0, Ts, "je", I.vvvo, "abei", 1.0, r, "ijab");          // :367
    contract(-1.0, Ts, "ma", I.ovoo, "mbij", 1.0, r, "ijab");          // :370
    contract(1.0, Ts, "mb", I.ovoo, "maij", 1.0, r, "ijab");           // :371
    Tensor t = T4d(no, no, nv, nv);
    for (long x = 0; x < r.size(); ++x) t.v[(size_t) x] = r.v[(size_t) x] / D_ijab.v[(size_t) x];// :374-383
    return t;
}

// cc.cpp:390-405, summed in the reference's loop order (i, a, j, b)
static double ccsd_energy(const Tensor &ts, const Tensor &td, const Tensor &oovv, const Tensor &F_ia) {
    const long no = ts.extent(0), nv = ts.extent(1);
    double e = 0.0;
    for (long i = 0; i < no; ++i)
        for (long a = 0; a < nv; ++a) {
            e += F_ia(i, a) * ts(i, a);
            for (long j = 0; j < no; ++j)
                for (long b = 0; b < nv; ++b)
                    e += 0.25 * oovv(i, j, a, b) * td(i, j, a, b) + 0.5 * oovv(i, j, a, b) * ts(i, a) * ts(j, b);
        }
    return e;
}

// cc.cpp:454-560, literal: materialises D, dT, tempTc, cT as o^3 v^3 tensors and issues
// the 18 contractions one by one.  Only usable for small o, v.
static double ccsd_t_literal(const Tensor &T1, const Tensor &T2, const Ints &I, const Moes &mo) {
    const long no = T1.extent(0), nv = T1.extent(1);
    const std::vector<long> e6{no, no, no, nv, nv, nv};
    Tensor D(e6), dT(e6), Tc(e6);
#pragma omp parallel for collapse(3) schedule(static)
    for (long i = 0; i < no; ++i)
        for (long j = 0; j < no; ++j)
            for (long k = 0; k < no; ++k)
                for (long a = 0; a < nv; ++a)
                    for (long b = 0; b < nv; ++b)
                        for (long c = 0; c < nv; ++c) {
                            const double d = mo.F_ii(i, i) + mo.F_ii(j, j) + mo.F_ii(k, k) - mo.F_aa(a, a) -
                                             mo.F_aa(b, b) - mo.F_aa(c, c);// :102-123
                            D(i, j, k, a, b, c) = d;
                            dT(i, j, k, a, b, c) +=                        // :472-496
                                    (T1(i, a) * I.oovv(j, k, b, c) - T1(i, b) * I.oovv(j, k, a, c) -
                                     T1(i, c) * I.oovv(j, k, b, a) - T1(j, a) * I.oovv(i, k, b, c) +
                                     T1(j, b) * I.oovv(i, k, a, c) + T1(j, c) * I.oovv(i, k, b, a) -
                                     T1(k, a) * I.oovv(j, i, b, c) + T1(k, b) * I.oovv(j, i, a, c) +
                                     T1(k, c) * I.oovv(j, i, b, a)) /
                                    d;
                        }
    contract(1.0, T2, "jkae", I.vovv, "eibc", 1.0, Tc, "ijkabc"); // :501
    contract(-1.0, T2, "jkbe", I.vovv, "eiac", 1.0, Tc, "ijkabc");// :502
    contract(-1.0, T2, "jkce", I.vovv, "eiba", 1.0, Tc, "ijkabc");// :503
    contract(-1.0, T2, "ikae", I.vovv, "ejbc", 1.0, Tc, "ijkabc");// :505
    contract(1.0, T2, "ikbe", I.vovv, "ejac", 1.0, Tc, "ijkabc"); // :506
    contract(1.0, T2, "ikce", I.vovv, "ejba", 1.0, Tc, "ijkabc"); // :507
    contract(-1.0, T2, "jiae", I.vovv, "ekbc", 1.0, Tc, "ijkabc");// :509
    contract(1.0, T2, "jibe", I.vovv, "ekac", 1.0, Tc, "ijkabc"); // :510
    contract(1.0, T2, "jice", I.vovv, "ekba", 1.0, Tc, "ijkabc"); // :511
    contract(-1.0, T2, "imbc", I.ovoo, "majk", 1.0, Tc, "ijkabc");// :513
    contract(1.0, T2, "imac", I.ovoo, "mbjk", 1.0, Tc, "ijkabc"); // :514
    contract(1.0, T2, "imba", I.ovoo, "mcjk", 1.0, Tc, "ijkabc"); // :515
    contract(1.0, T2, "jmbc", I.ovoo, "maik", 1.0, Tc, "ijkabc"); // :517
    contract(-1.0, T2, "jmac", I.ovoo, "mbik", 1.0, Tc, "ijkabc");// :518
    contract(-1.0, T2, "jmba", I.ovoo, "mcik", 1.0, Tc, "ijkabc");// :519
    contract(1.0, T2, "kmbc", I.ovoo, "maji", 1.0, Tc, "ijkabc"); // :521
    contract(-1.0, T2, "kmac", I.ovoo, "mbji", 1.0, Tc, "ijkabc");// :522
    contract(-1.0, T2, "kmba", I.ovoo, "mcji", 1.0, Tc, "ijkabc");// :523
    double val = 0.0;
    for (long x = 0; x < Tc.size(); ++x) {
        const double cT = Tc.v[(size_t) x] / D.v[(size_t) x];               // :525-538
        val += (cT * D.v[(size_t) x]) * (cT + dT.v[(size_t) x]);            // :541-556
    }
    return val / 36;
}

// Same quantity evaluated one occupied triple at a time (SURVEY.md section 7 restatement):
// X = A.B with K = 3(v+o), W = X - X[b,a,c] - X[c,b,a], no o^3 v^3 storage.  Valid for
// arbitrary inputs (no symmetry assumed); used as the CPU path for sizes where the literal
// algorithm does not fit in memory.  triples = flat (i,j,k) list; each weighted by `weight`.
static double ccsd_t_blocked(const Tensor &T1, const Tensor &T2, const Tensor &oovv, const Tensor &vovv,
                             const Tensor &ovoo, const double *eo, const double *ev, const int *triples, long ntrip,
                             double weight) {
    const long o = T1.extent(0), v = T1.extent(1), K = 3 * (v + o), N = v * v;
    std::vector<double> A((size_t) (v * K)), B((size_t) (K * N)), X((size_t) (v * N));
    double total = 0.0;
    for (long t = 0; t < ntrip; ++t) {
        const long i = triples[3 * t], j = triples[3 * t + 1], k = triples[3 * t + 2];
#pragma omp parallel for schedule(static)
        for (long a = 0; a < v; ++a) {
            double *r = &A[(size_t) (a * K)];
            for (long e = 0; e < v; ++e) {
                r[e] = T2(j, k, a, e);
                r[v + e] = -T2(i, k, a, e);
                r[2 * v + e] = -T2(j, i, a, e);
            }
            for (long m = 0; m < o; ++m) {
                r[3 * v + m] = -ovoo(m, a, j, k);
                r[3 * v + o + m] = ovoo(m, a, i, k);
                r[3 * v + 2 * o + m] = ovoo(m, a, j, i);
            }
        }
        const long occ3[3] = {i, j, k};
#pragma omp parallel for schedule(static)
        for (long e = 0; e < v; ++e)
            for (int s = 0; s < 3; ++s)
                std::memcpy(&B[(size_t) ((s * v + e) * N)], &vovv.v[(size_t) vovv.off(e, occ3[s], 0, 0)],
                            (size_t) N * sizeof(double));
#pragma omp parallel for schedule(static)
        for (long m = 0; m < o; ++m)
            for (int s = 0; s < 3; ++s)
                std::memcpy(&B[(size_t) ((3 * v + s * o + m) * N)], &T2.v[(size_t) T2.off(occ3[s], m, 0, 0)],
                            (size_t) N * sizeof(double));
        gemm_acc(v, N, K, A.data(), B.data(), X.data());
        const double dijk = eo[i] + eo[j] + eo[k];
        double et = 0.0;
#pragma omp parallel for reduction(+ : et) schedule(static)
        for (long a = 0; a < v; ++a)
            for (long b = 0; b < v; ++b)
                for (long c = 0; c < v; ++c) {
                    auto Xa = [&](long p, long q, long r) { return X[(size_t) ((p * v + q) * v + r)]; };
                    auto Y = [&](long p, long q, long r) {
                        return T1(i, p) * oovv(j, k, q, r) - T1(j, p) * oovv(i, k, q, r) - T1(k, p) * oovv(j, i, q, r);
                    };
                    const double W = Xa(a, b, c) - Xa(b, a, c) - Xa(c, b, a);
                    const double V = Y(a, b, c) - Y(b, a, c) - Y(c, b, a);
                    const double d = dijk - ev[a] - ev[b] - ev[c];
                    et += W * (W + V) / d;
                }
        total += weight * et;
    }
    return total / 36;
}

// Energy sweep of one sorted triple from the b<c-packed X = A.B (CPU baseline only): the same
// W / V / D arithmetic as ccsd_t_blocked restricted to a<b<c (needs antisymmetric inputs), all
// host threads, no copies of the big tensors.  Returns sum_{a<b<c} W (W+V) / D, i.e. the triple's
// full contribution to E(T).
static double t_energy_packed(const double *Xp, const double *T1, const double *oovv, const double *eo, const double *ev,
                              long o, long v, long i, long j, long k) {
    const long np = v * (v - 1) / 2;
    auto pidx = [v](long b, long c) { return b * v - b * (b + 1) / 2 + (c - b - 1); };
    const double dijk = eo[i] + eo[j] + eo[k];
    const double *Ojk = oovv + (j * o + k) * v * v, *Oik = oovv + (i * o + k) * v * v, *Oji = oovv + (j * o + i) * v * v;
    const double *ti = T1 + i * v, *tj = T1 + j * v, *tk = T1 + k * v;
    double total = 0.0;
#pragma omp parallel for reduction(+ : total) schedule(dynamic, 4)
    for (long a = 0; a < v - 2; ++a) {
        double acc = 0.0;
        for (long b = a + 1; b < v - 1; ++b) {
            const long pab = pidx(a, b);
            for (long c = b + 1; c < v; ++c) {
                const long pbc = pidx(b, c), pac = pidx(a, c);
                const double W = Xp[a * np + pbc] - Xp[b * np + pac] + Xp[c * np + pab];
                const double Ya = ti[a] * Ojk[b * v + c] - tj[a] * Oik[b * v + c] - tk[a] * Oji[b * v + c];
                const double Yb = ti[b] * Ojk[a * v + c] - tj[b] * Oik[a * v + c] - tk[b] * Oji[a * v + c];
                const double Yc = ti[c] * Ojk[a * v + b] - tj[c] * Oik[a * v + b] - tk[c] * Oji[a * v + b];
                acc += W * (W + (Ya - Yb + Yc)) / (dijk - ev[a] - ev[b] - ev[c]);
            }
        }
        total += acc;
    }
    return total;
}

}// namespace orc

// ---------------------------------------------------------------------------------
// flat C interface for ctypes (tests / bench cpu_baseline)
// ---------------------------------------------------------------------------------
using orc::Tensor;
typedef std::vector<long> Ext;

static orc::Ints wrap_ints(const double *const *p, long o, long v) {
    orc::Ints s;
    for (int k = 0; k < 11; ++k) {
        Ext e(4);
        for (int d = 0; d < 4; ++d) e[(size_t) d] = orc::kIntNames[k][d] == 'o' ? o : v;
        *orc::int_field(s, k) = p[k] ? Tensor(e, p[k]) : Tensor();
    }
    return s;
}
static orc::Moes wrap_moes(const double *foo, const double *fov, const double *fvv, long o, long v) {
    orc::Moes m;
    m.F_ii = Tensor(Ext{o, o}, foo);
    m.F_ia = Tensor(Ext{o, v}, fov);// fov == nullptr -> zeros (Q4: the only defined reading for UHF)
    m.F_aa = Tensor(Ext{v, v}, fvv);
    return m;
}
static void put(const Tensor &t, double *dst) {
    if (dst) std::memcpy(dst, t.data(), (size_t) t.size() * sizeof(double));
}

#define ORC_TRY try {
#define ORC_CATCH                                            \
    }                                                        \
    catch (const std::exception &e) {                        \
        std::fprintf(stderr, "oracle error: %s\n", e.what()); \
        return 1;                                            \
    }                                                        \
    return 0;

extern "C" {

int orc_contract(double alpha, const double *A, const long *extA, const char *ia, const double *B, const long *extB,
                 const char *ib, double beta, double *C, const long *extC, const char *ic) {
    ORC_TRY
    Tensor a(Ext(extA, extA + std::strlen(ia)), A), b(Ext(extB, extB + std::strlen(ib)), B),
            c(Ext(extC, extC + std::strlen(ic)), C);
    orc::contract(alpha, a, ia, b, ib, beta, c, ic);
    put(c, C);
    ORC_CATCH
}

int orc_transform_ao_mo(const double *ao, const double *C1, const double *C2, int n, double *out) {
    ORC_TRY
    put(orc::transform_ao_mo(Tensor(Ext{n, n, n, n}, ao), Tensor(Ext{n, n}, C1), Tensor(Ext{n, n}, C2)), out);
    ORC_CATCH
}

int orc_transform_to_so(const double *aa, const double *bb, const double *ab, int n, double *out) {
    ORC_TRY
    const Ext e{n, n, n, n};
    put(orc::transform_to_so(Tensor(e, aa), Tensor(e, bb), Tensor(e, ab)), out);
    ORC_CATCH
}

int orc_slice_ints(const double *so, int n2, int no, int nv, const char *spec, double *out) {
    ORC_TRY
    put(orc::slice_ints(Tensor(Ext{n2, n2, n2, n2}, so), no, nv, spec), out);
    ORC_CATCH
}

int orc_make_so_moes(const double *ea, const double *eb, int nao, double *out) {
    ORC_TRY
    put(orc::make_so_moes(ea, eb, nao), out);
    ORC_CATCH
}

int orc_make_denom(const double *F_spin, int n2, int no, int nv, double *out) {
    ORC_TRY
    put(orc::make_denom(Tensor(Ext{n2, n2}, F_spin), no, nv), out);
    ORC_CATCH
}

int orc_run_mp2(const double *oovv, const double *denom, int no, int nv, double *T, double *E) {
    ORC_TRY
    const Ext e{no, no, nv, nv};
    Tensor t;
    *E = orc::run_mp2(Tensor(e, oovv), Tensor(e, denom), t);
    put(t, T);
    ORC_CATCH
}

int orc_make_fock(const double *e1, const double *e2, int n1, int n2, double *out) {
    ORC_TRY
    put(orc::make_fock(e1, e2, n1, n2), out);
    ORC_CATCH
}

int orc_make_tau(const double *Ts, const double *Td, int no, int nv, int bar, double *out) {
    ORC_TRY
    Tensor ts(Ext{no, nv}, Ts), td(Ext{no, no, nv, nv}, Td);
    put(bar ? orc::make_tau_bar(ts, td) : orc::make_tau(ts, td), out);
    ORC_CATCH
}

int orc_multiply_Ts(const double *Ts, int no, int nv, double *out) {
    ORC_TRY
    put(orc::multiply_Ts(Tensor(Ext{no, nv}, Ts)), out);
    ORC_CATCH
}

// ints: 11 pointers in the order of integrals.h:11-23; out: F_ae, F_mi, F_me, W_mnij, W_mbej, W_abef
int orc_update_intermediates(const double *Ts, const double *Td, const double *const *ints, const double *foo,
                             const double *fov, const double *fvv, int no, int nv, double *const *out) {
    ORC_TRY
    orc::Inter X = orc::update_intermediates(Tensor(Ext{no, nv}, Ts), Tensor(Ext{no, no, nv, nv}, Td),
                                             wrap_ints(ints, no, nv), wrap_moes(foo, fov, fvv, no, nv));
    put(X.F_ae, out[0]), put(X.F_mi, out[1]), put(X.F_me, out[2]);
    put(X.W_mnij, out[3]), put(X.W_mbej, out[4]), put(X.W_abef, out[5]);
    ORC_CATCH
}

static orc::Inter wrap_inter(const double *const *p, long o, long v) {
    orc::Inter X;
    X.F_ae = Tensor(Ext{v, v}, p[0]), X.F_mi = Tensor(Ext{o, o}, p[1]), X.F_me = Tensor(Ext{o, v}, p[2]);
    X.W_mnij = Tensor(Ext{o, o, o, o}, p[3]), X.W_mbej = Tensor(Ext{o, v, v, o}, p[4]);
    X.W_abef = Tensor(Ext{v, v, v, v}, p[5]);
    return X;
}

int orc_make_T1(const double *Ts, const double *Td, const double *const *ints, const double *const *inter,
                const double *foo, const double *fov, const double *fvv, int no, int nv, double *out) {
    ORC_TRY
    orc::Moes mo = wrap_moes(foo, fov, fvv, no, nv);
    put(orc::make_T1(Tensor(Ext{no, nv}, Ts), Tensor(Ext{no, no, nv, nv}, Td), wrap_ints(ints, no, nv),
                     wrap_inter(inter, no, nv), orc::make_D_ia(mo), mo),
        out);
    ORC_CATCH
}

int orc_make_T2(const double *Ts, const double *Td, const double *const *ints, const double *const *inter,
                const double *foo, const double *fvv, int no, int nv, double *out) {
    ORC_TRY
    orc::Moes mo = wrap_moes(foo, nullptr, fvv, no, nv);
    put(orc::make_T2(Tensor(Ext{no, nv}, Ts), Tensor(Ext{no, no, nv, nv}, Td), wrap_ints(ints, no, nv),
                     wrap_inter(inter, no, nv), orc::make_D_ijab(mo)),
        out);
    ORC_CATCH
}

int orc_ccsd_energy(const double *Ts, const double *Td, const double *oovv, const double *fov, int no, int nv,
                    double *E) {
    ORC_TRY
    *E = orc::ccsd_energy(Tensor(Ext{no, nv}, Ts), Tensor(Ext{no, no, nv, nv}, Td), Tensor(Ext{no, no, nv, nv}, oovv),
                          Tensor(Ext{no, nv}, fov));
    ORC_CATCH
}

// cc.cpp:409-451.  T1 starts at zero, T2 at the supplied guess (the MP2 amplitudes); the
// energy is evaluated before each update and the loop stops when |dE| < conv (fabs; Q3).
// e_hist (length maxiter) receives the energy printed at each iteration.
int orc_ccsd(const double *const *ints, const double *foo, const double *fov, const double *fvv, const double *T2_guess,
             int no, int nv, int maxiter, double conv, double *T1_out, double *T2_out, double *E_out, int *iters_out,
             double *e_hist) {
    ORC_TRY
    orc::Ints I = wrap_ints(ints, no, nv);
    orc::Moes mo = wrap_moes(foo, fov, fvv, no, nv);
    Tensor D1 = orc::make_D_ia(mo), D2 = orc::make_D_ijab(mo);
    Tensor T1(Ext{no, nv}), T2(Ext{no, no, nv, nv}, T2_guess);
    double e_last = 0.0, e = 0.0;
    int it = 0;
    for (; it < maxiter; ++it) {
        e = orc::ccsd_energy(T1, T2, I.oovv, mo.F_ia);
        const double de = e - e_last;
        e_last = e;
        if (e_hist) e_hist[it] = e;
        if (std::fabs(de) < conv) break;
        orc::Inter X = orc::update_intermediates(T1, T2, I, mo);// :442
        T1 = orc::make_T1(T1, T2, I, X, D1, mo);                // :445 (new T1 ...)
        T2 = orc::make_T2(T1, T2, I, X, D2);                    // :446 (... feeds make_T2 with old intermediates, Q1)
    }
    put(T1, T1_out), put(T2, T2_out);
    *E_out = e;
    if (iters_out) *iters_out = it;
    ORC_CATCH
}

int orc_ccsd_t(const double *T1, const double *T2, const double *oovv, const double *vovv, const double *ovoo,
               const double *foo, const double *fvv, int no, int nv, double *E) {
    ORC_TRY
    orc::Ints I;
    I.oovv = Tensor(Ext{no, no, nv, nv}, oovv), I.vovv = Tensor(Ext{nv, no, nv, nv}, vovv);
    I.ovoo = Tensor(Ext{no, nv, no, no}, ovoo);
    *E = orc::ccsd_t_literal(Tensor(Ext{no, nv}, T1), Tensor(Ext{no, no, nv, nv}, T2), I,
                             wrap_moes(foo, nullptr, fvv, no, nv));
    ORC_CATCH
}

int orc_t_energy_packed(const double *Xp, const double *T1, const double *oovv, const double *eps_o, const double *eps_v,
                        int no, int nv, int i, int j, int k, double *E) {
    ORC_TRY
    *E = orc::t_energy_packed(Xp, T1, oovv, eps_o, eps_v, no, nv, i, j, k);
    ORC_CATCH
}

int orc_ccsd_t_blocked(const double *T1, const double *T2, const double *oovv, const double *vovv, const double *ovoo,
                       const double *eps_o, const double *eps_v, int no, int nv, const int *triples, long ntrip,
                       double weight, double *E) {
    ORC_TRY
    *E = orc::ccsd_t_blocked(Tensor(Ext{no, nv}, T1), Tensor(Ext{no, no, nv, nv}, T2), Tensor(Ext{no, no, nv, nv}, oovv),
                             Tensor(Ext{nv, no, nv, nv}, vovv), Tensor(Ext{no, nv, no, no}, ovoo), eps_o, eps_v,
                             triples, ntrip, weight);
    ORC_CATCH
}

}// extern "C"
