// contract.cu -- device replacement of btas::contract, the reference's only arithmetic
// primitive (71 call sites: integrals.cpp:121-129, cc.cpp:194-371, 501-523):
//     C[ic] = beta*C[ic] + alpha * sum_{labels in A and B only} A[ia]*B[ib]
// Each call is mapped onto ONE DMMA GEMM.  Operands whose index order already factors into
// (free block, contracted block) in either order are consumed in place through the GEMM's
// transposed-operand variants; a leading index shared by C and one operand is peeled off as
// the GEMM batch dimension; only operands with interleaved index groups are re-laid out
// (bandwidth-bound permute kernel), and the result is permuted back only if C's own index
// order interleaves the two free groups.
#include <algorithm>
#include <cstring>
#include <string>

#include "contract.cuh"
#include "dgemm.cuh"
#include "tensor_ops.cuh"

namespace qcp2 {

namespace {

struct Spec {
    std::string a, b, c;               // label strings (after peeling)
    long long ea[8], eb[8], ec[8];     // extents
};

long long ext_of(const std::string &s, const long long *e, char l) {
    const size_t p = s.find(l);
    return e[p];
}
long long prod(const std::string &labels, const std::string &s, const long long *e) {
    long long n = 1;
    for (char l : labels) n *= ext_of(s, e, l);
    return n;
}
long long total(const std::string &s, const long long *e) {
    long long n = 1;
    for (size_t k = 0; k < s.size(); ++k) n *= e[k];
    return n;
}

struct Plan {
    std::string M, N, K;// canonical label orders
    bool a_direct = false, a_kc = true;
    bool b_direct = false, b_nc = true;
    bool c_direct = false, c_swapped = false;
    double cost = 0.0;
};

Plan make_plan(const Spec &s, const std::string &M, const std::string &N, const std::string &K, bool scatter_ok) {
    Plan p;
    p.M = M, p.N = N, p.K = K;
    if (s.a == M + K) p.a_direct = true, p.a_kc = true;
    else if (s.a == K + M) p.a_direct = true, p.a_kc = false;
    if (s.b == K + N) p.b_direct = true, p.b_nc = true;
    else if (s.b == N + K) p.b_direct = true, p.b_nc = false;
    if (s.c == M + N) p.c_direct = true, p.c_swapped = false;
    else if (s.c == N + M) p.c_direct = true, p.c_swapped = true;
    // a result written through offset tables costs a little (8-byte stores); without them a dense product + re-layout pass
    p.cost = (p.a_direct ? 0.0 : 2.0 * (double) total(s.a, s.ea)) + (p.b_direct ? 0.0 : 2.0 * (double) total(s.b, s.eb)) +
             (p.c_direct ? 0.0 : (scatter_ok ? 0.25 : 3.0) * (double) total(s.c, s.ec));
    return p;
}

Plan best_plan(const Spec &s, bool scatter_ok) {
    std::string M, N, Ka, Kb;
    for (char l : s.c) {
        const bool ina = s.a.find(l) != std::string::npos, inb = s.b.find(l) != std::string::npos;
        QCP2_REQUIRE(ina != inb, "contract: every output label must occur in exactly one operand");
        (ina ? M : N).push_back(l);
    }
    for (char l : s.a)
        if (s.c.find(l) == std::string::npos) {
            QCP2_REQUIRE(s.b.find(l) != std::string::npos, "contract: dangling label in A");
            Ka.push_back(l);
        }
    for (char l : s.b)
        if (s.c.find(l) == std::string::npos) {
            QCP2_REQUIRE(s.a.find(l) != std::string::npos, "contract: dangling label in B");
            Kb.push_back(l);
        }
    QCP2_REQUIRE(Ka.size() == Kb.size(), "contract: contracted label sets differ");
    for (char l : Ka) QCP2_REQUIRE(ext_of(s.a, s.ea, l) == ext_of(s.b, s.eb, l), "contract: contracted extents differ");
    for (char l : M) QCP2_REQUIRE(ext_of(s.a, s.ea, l) == ext_of(s.c, s.ec, l), "contract: free extents differ (A)");
    for (char l : N) QCP2_REQUIRE(ext_of(s.b, s.eb, l) == ext_of(s.c, s.ec, l), "contract: free extents differ (B)");
    // candidate orders of the free labels: as they appear in C, or as they appear in their own operand (then the
    // operand is consumed in place and C is written through the GEMM's offset tables -- cheaper than re-laying a large
    // operand out); contracted labels in A's or B's order
    std::string Ma, Nb;
    for (char l : s.a)
        if (s.c.find(l) != std::string::npos) Ma.push_back(l);
    for (char l : s.b)
        if (s.c.find(l) != std::string::npos) Nb.push_back(l);
    Plan best = make_plan(s, M, N, Ka, scatter_ok);
    const std::string *Ms[2] = {&M, &Ma}, *Ns[2] = {&N, &Nb}, *Ks[2] = {&Ka, &Kb};
    for (int im = 0; im < 2; ++im)
        for (int in = 0; in < 2; ++in)
            for (int ik = 0; ik < 2; ++ik) {
                if ((im && Ma == M) || (in && Nb == N) || (ik && Kb == Ka)) continue;
                Plan p = make_plan(s, *Ms[im], *Ns[in], *Ks[ik], scatter_ok);
                if (p.cost < best.cost) best = p;
            }
    return best;
}

// re-lay `src` (labels `from`) out as a dense tensor with labels `to`
void relayout(Ctx &ctx, double *dst, const double *src, const std::string &from, const long long *efrom,
              const std::string &to, double alpha, double beta) {
    long long istr_from[8];
    long long s = 1;
    for (int k = (int) from.size() - 1; k >= 0; --k) istr_from[k] = s, s *= efrom[k];
    long long oext[8], istr[8];
    for (size_t k = 0; k < to.size(); ++k) {
        const size_t p = from.find(to[k]);
        oext[k] = efrom[p], istr[k] = istr_from[p];
    }
    permute_axpby(ctx, dst, src, (int) to.size(), oext, istr, alpha, beta);
}

// operand re-layout of a (possibly batched) tensor, served from the context's cache when the source
// is registered as constant; returns the buffer to feed to the GEMM
const double *operand_relayout(Ctx &ctx, DBuf &tmp, const double *src, long long stride, int nb, const std::string &from,
                               const long long *efrom, const std::string &to) {
    const long long n = total(from, efrom);
    PermCache *pc = ctx.perm_cache;
    const bool cacheable = pc && pc->is_constant(src);
    std::string sig;
    if (cacheable) {
        sig = from + ">" + to + ":" + std::to_string(nb) + ":" + std::to_string(stride);
        for (size_t k = 0; k < from.size(); ++k) sig += "," + std::to_string(efrom[k]);
        for (const PermCache::Entry &e : pc->entries)
            if (e.src == src && e.sig == sig) return e.buf;
    }
    const bool store = cacheable && !pc->frozen;
    double *dst;
    if (store) dst = ctx.alloc((size_t) (n * nb));
    else tmp = DBuf(ctx, (size_t) (n * nb)), dst = tmp.p;
    for (int z = 0; z < nb; ++z) relayout(ctx, dst + z * n, src + z * stride, from, efrom, to, 1.0, 0.0);
    if (store) pc->entries.push_back(PermCache::Entry{src, sig, dst});
    return dst;
}

// Offset tables of a result whose index order interleaves the free labels of A and B:
// &C[...] = C + row_off[m] + col_off[n], m / n the composite M / N indices in the plan's label order.
// One device array [M + N], cached in the context under the call's signature.
const long long *offset_tables(Ctx &ctx, const Spec &s, const Plan &p, long long M, long long N) {
    std::string sig = s.c + "|" + p.M + "|" + p.N;
    for (size_t k = 0; k < s.c.size(); ++k) sig += "," + std::to_string(s.ec[k]);
    for (const Ctx::OffsetTable &t : ctx.offset_tables)
        if (t.sig == sig) return t.dev;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx.stream, &cap);
    if (cap != cudaStreamCaptureStatusNone) return nullptr;// cannot allocate / copy synchronously inside a capture
    long long cstr[8], st = 1;
    for (int k = (int) s.c.size() - 1; k >= 0; --k) cstr[k] = st, st *= s.ec[k];
    std::vector<long long> tab((size_t) (M + N));
    auto fill = [&](const std::string &labels, long long count, long long *out) {
        for (long long lin = 0; lin < count; ++lin) {
            long long rem = lin, off = 0;
            for (int q = (int) labels.size() - 1; q >= 0; --q) {
                const size_t pos = s.c.find(labels[q]);
                const long long e = s.ec[pos];
                off += (rem % e) * cstr[pos];
                rem /= e;
            }
            out[lin] = off;
        }
    };
    fill(p.M, M, tab.data());
    fill(p.N, N, tab.data() + M);
    long long *dev = nullptr;
    QCP2_CUDA(cudaMalloc(&dev, tab.size() * sizeof(long long)));
    QCP2_CUDA(cudaMemcpyAsync(dev, tab.data(), tab.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
    QCP2_CUDA(cudaStreamSynchronize(ctx.stream));// `tab` is pageable and goes out of scope
    ctx.offset_tables.push_back(Ctx::OffsetTable{sig, dev});
    return dev;
}

void run(Ctx &ctx, double alpha, const double *A, const double *B, double beta, double *C, const Spec &s,
         const Plan &p, int batch, long long sA, long long sB, long long sC, const double *cin) {
    const long long M = prod(p.M, s.c, s.ec), N = prod(p.N, s.c, s.ec), K = prod(p.K, s.a, s.ea);
    QCP2_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "contract: GEMM extent exceeds 2^31");
    DBuf ta, tb, tc;
    const double *Ag = A, *Bg = B;
    long long sAg = sA, sBg = sB;
    bool a_kc = p.a_kc, b_nc = p.b_nc;
    if (!p.a_direct) {// -> [M][K]
        const long long n = total(s.a, s.ea);
        Ag = operand_relayout(ctx, ta, A, sA, sA ? batch : 1, s.a, s.ea, p.M + p.K);
        a_kc = true, sAg = sA ? n : 0;
    }
    if (!p.b_direct) {// -> [K][N]
        const long long n = total(s.b, s.eb);
        Bg = operand_relayout(ctx, tb, B, sB, sB ? batch : 1, s.b, s.eb, p.K + p.N);
        b_nc = true, sBg = sB ? n : 0;
    }
    GemmArgs g;
    g.batch = batch;
    g.alpha = alpha;
    double *Cg = C;
    long long sCg = sC;
    bool swapped = p.c_direct && p.c_swapped;
    bool scatter = false;
    if (!p.c_direct && !getenv("QCP2_NO_SCATTER_EPILOGUE") && !getenv("QCP2_NO_WS_GEMM")) {
        // the result's index order interleaves the free labels of the two operands (cc.cpp:230, 246-249): the GEMM
        // writes it in place through row / column offset tables instead of a dense product + re-layout pass
        GemmArgs probe;
        probe.A = Ag, probe.lda = a_kc ? K : M, probe.strideA = sAg, probe.B = Bg, probe.ldb = b_nc ? N : K, probe.strideB = sBg;
        probe.K = (int) K, probe.batch = batch;
        if (gemm_takes_tables(ctx, probe)) {
            const long long *tab = offset_tables(ctx, s, p, M, N);
            if (tab) {
                scatter = true;
                g.row_off = tab, g.col_off = tab + M;
                g.beta = beta;
                if (cin && beta != 0.0) g.Cin = cin, g.strideCin = sC;
            }
        }
    }
    if (scatter) {
    } else if (!p.c_direct) {
        const long long n = total(s.c, s.ec);
        if (cin && cin != C && beta != 0.0) axpby(ctx, C, cin, 1.0, 0.0, n * batch);// the result goes through a re-layout: seed C with cin
        tc = DBuf(ctx, (size_t) (n * batch));
        Cg = tc.p, sCg = n, g.beta = 0.0;
    } else {
        g.beta = beta;
        if (cin && beta != 0.0) g.Cin = cin, g.strideCin = sC;// same layout as C
    }
    if (!swapped) {
        g.M = (int) M, g.N = (int) N, g.K = (int) K;
        g.A = Ag, g.lda = a_kc ? K : M, g.strideA = sAg;
        g.B = Bg, g.ldb = b_nc ? N : K, g.strideB = sBg;
        g.C = Cg, g.ldc = N, g.strideC = sCg, g.ldcin = N;
        gemm(ctx, g, a_kc, b_nc);
    } else {// C^T[N][M] = B^T[N][K] . A^T[K][M]
        g.M = (int) N, g.N = (int) M, g.K = (int) K;
        g.A = Bg, g.lda = b_nc ? N : K, g.strideA = sBg;
        g.B = Ag, g.ldb = a_kc ? K : M, g.strideB = sAg;
        g.C = Cg, g.ldc = M, g.strideC = sCg, g.ldcin = M;
        gemm(ctx, g, /*a_kc=*/!b_nc, /*b_nc=*/!a_kc);
    }
    if (!p.c_direct && !scatter) {
        const std::string mn = p.M + p.N;
        long long emn[8];
        for (size_t k = 0; k < mn.size(); ++k) emn[k] = ext_of(s.c, s.ec, mn[k]);
        const long long n = total(s.c, s.ec);
        for (int z = 0; z < batch; ++z) relayout(ctx, C + z * sC, tc.p + z * n, mn, emn, s.c, 1.0, beta);
    }
}

}// namespace

void contract(Ctx &ctx, double alpha, const TView &A, const char *ia, const TView &B, const char *ib, double beta,
              const TView &C, const char *ic, const double *cin) {
    Spec s;
    s.a = ia, s.b = ib, s.c = ic;
    QCP2_REQUIRE((int) s.a.size() == A.rank && (int) s.b.size() == B.rank && (int) s.c.size() == C.rank,
                 "contract: label count does not match tensor rank");
    for (int k = 0; k < A.rank; ++k) s.ea[k] = A.ext[k];
    for (int k = 0; k < B.rank; ++k) s.eb[k] = B.ext[k];
    for (int k = 0; k < C.rank; ++k) s.ec[k] = C.ext[k];
    if (C.size() == 0) return;
    // is the GEMM's offset-table epilogue available to this call?  (the planner prices interleaved results with it)
    static const bool no_scatter = getenv("QCP2_NO_SCATTER_EPILOGUE") != nullptr || getenv("QCP2_NO_WS_GEMM") != nullptr;
    const bool scatter_ok = !no_scatter && ctx.ws_counters != nullptr && !(ctx.shard) &&
                            (reinterpret_cast<uintptr_t>(A.p) & 15) == 0 && (reinterpret_cast<uintptr_t>(B.p) & 15) == 0;
    Plan p0 = best_plan(s, scatter_ok);
    // try peeling the leading label of C as a batch dimension
    bool peel = false;
    Spec sp = s;
    Plan pp;
    long long sA = 0, sB = 0, sC = 0;
    int batch = 1;
    if (p0.cost > 0.0 && s.c.size() >= 2) {
        const char l = s.c[0];
        const bool lead_a = !s.a.empty() && s.a[0] == l, lead_b = !s.b.empty() && s.b[0] == l;
        if ((lead_a || lead_b) && s.ec[0] <= 65535) {
            sp.c = s.c.substr(1);
            for (size_t k = 0; k + 1 < s.c.size(); ++k) sp.ec[k] = s.ec[k + 1];
            if (lead_a) {
                sp.a = s.a.substr(1);
                for (size_t k = 0; k + 1 < s.a.size(); ++k) sp.ea[k] = s.ea[k + 1];
                sA = total(sp.a, sp.ea);
            } else {
                sp.b = s.b.substr(1);
                for (size_t k = 0; k + 1 < s.b.size(); ++k) sp.eb[k] = s.eb[k + 1];
                sB = total(sp.b, sp.eb);
            }
            sC = total(sp.c, sp.ec);
            batch = (int) s.ec[0];
            pp = best_plan(sp, scatter_ok);
            // cost of the peeled plan counts every batch member of the batched tensors
            const double scaled = (pp.a_direct ? 0.0 : 2.0 * (double) total(sp.a, sp.ea) * (sA ? batch : 1)) +
                                  (pp.b_direct ? 0.0 : 2.0 * (double) total(sp.b, sp.eb) * (sB ? batch : 1)) +
                                  (pp.c_direct ? 0.0 : (scatter_ok ? 0.25 : 3.0) * (double) total(sp.c, sp.ec) * batch);
            if (scaled < p0.cost) peel = true;
        }
    }
    const Plan &chosen = peel ? pp : p0;
    if (peel) run(ctx, alpha, A.p, B.p, beta, C.p, sp, pp, batch, sA, sB, sC, cin);
    else run(ctx, alpha, A.p, B.p, beta, C.p, s, p0, 1, 0, 0, 0, cin);
}

}// namespace qcp2
