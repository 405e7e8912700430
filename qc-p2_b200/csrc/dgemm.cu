// dgemm.cu -- alignment analysis, tile selection and split-K for the DMMA GEMM.
#include <cstdio>
#include <cstdlib>

#include "comm.cuh"
#include "dgemm.cuh"
#include "dgemm_ws.cuh"
#include "tensor_ops.cuh"

namespace qcp2 {

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// sums the split-K partials in split order (deterministic): C = alpha * sum_s P[s] + beta * Cin
__global__ void splitk_reduce_kernel(const double *__restrict__ P, int splits, long long mn, int N, double *__restrict__ C,
                                     long long ldc, long long strideC, const double *__restrict__ Cin, long long ldcin,
                                     long long strideCin, double alpha, double beta, int batch) {
    const long long total = mn * batch;
    for (long long lin = (long long) blockIdx.x * blockDim.x + threadIdx.x; lin < total; lin += (long long) gridDim.x * blockDim.x) {
        const long long z = lin / mn, e = lin - z * mn;
        const long long row = e / N, col = e - row * N;
        double s = 0.0;
        for (int q = 0; q < splits; ++q) s += P[(z * splits + q) * mn + e];
        s *= alpha;
        if (beta != 0.0) s += beta * Cin[z * strideCin + row * ldcin + col];
        C[z * strideC + row * ldc + col] = s;
    }
}

struct TileChoice {
    int bm, bn;
    double eff;// rough per-CTA tensor-pipe efficiency of the tile shape
};

// ---- warp-specialised TMA path (dgemm_ws.cuh) ---------------------------------------------
namespace {

bool ws_eligible(const Ctx &ctx, const GemmArgs &g) {
    static const bool off = getenv("QCP2_NO_WS_GEMM") != nullptr;
    if (off || g.seg != nullptr || !ws_tma_available() || ctx.ws_counters == nullptr) return false;
    if (g.K < 1) return false;
    const bool batched = g.batch > 1;
    if (!aligned16(g.A) || (g.lda & 1) || (batched && (g.strideA & 1))) return false;
    if (!aligned16(g.B) || (g.ldb & 1) || (batched && (g.strideB & 1))) return false;
    if (g.Cin && g.Cin != g.C && !g.row_off && (g.ldcin != g.ldc)) return false;// one addressing for C and Cin
    return true;
}

void ws_dispatch(Ctx &ctx, int shape, const CUtensorMap &ta, const CUtensorMap &tb, const WsGemmArgs &w, bool a_kc, bool b_nc) {
    switch (shape) {
        case WS_80x256: ws_launch_80x256(ctx, ta, tb, w, a_kc, b_nc); break;
        case WS_48x256: ws_launch_48x256(ctx, ta, tb, w, a_kc, b_nc); break;
        case WS_16x256: ws_launch_16x256(ctx, ta, tb, w, a_kc, b_nc); break;
        case WS_64x64: ws_launch_64x64(ctx, ta, tb, w, a_kc, b_nc); break;
        case WS_16x192: ws_launch_16x192(ctx, ta, tb, w, a_kc, b_nc); break;
        case WS_32x256: ws_launch_32x256(ctx, ta, tb, w, a_kc, b_nc); break;
        case WS_96x128: ws_launch_96x128(ctx, ta, tb, w, a_kc, b_nc); break;
        default: ws_launch_112x128(ctx, ta, tb, w, a_kc, b_nc); break;
    }
}

// runs g on the warp-specialised kernel: tile `shape`, optionally as the transposed problem C^T = B^T A^T
// (so that a skinny N takes the skinny-M tiles), split-K factor `split`
void ws_run(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int shape, bool swap, int split) {
    const WsShapeInfo &sh = kWsShapes[shape];
    WsGemmArgs w;
    const double *A = g.A, *B = g.B;
    long long lda = g.lda, ldb = g.ldb, sA = g.strideA, sB = g.strideB;
    w.M = g.M, w.N = g.N, w.K = g.K;
    w.c_rs = g.ldc, w.c_cs = 1;
    w.row_off = g.row_off, w.col_off = g.col_off;
    if (swap) {
        std::swap(A, B), std::swap(lda, ldb), std::swap(sA, sB);
        std::swap(w.M, w.N), std::swap(w.c_rs, w.c_cs), std::swap(w.row_off, w.col_off);
        const bool t = a_kc;
        a_kc = !b_nc, b_nc = !t;
    }
    w.batch = g.batch, w.a_batched = (g.batch > 1 && sA != 0), w.b_batched = (g.batch > 1 && sB != 0);
    w.alpha = g.alpha, w.beta = g.beta;
    w.C = g.C, w.Cin = g.Cin, w.strideC = g.strideC, w.strideCin = g.Cin ? g.strideCin : g.strideC;
    // A: [M][K] (K contiguous) -> {K, M}, box {16, BM};  [K][M] -> {M, K}, boxes {16, 16}
    const CUtensorMap ta = a_kc ? ws_make_map(A, w.K, w.M, lda, g.batch, sA, sh.bm) : ws_make_map(A, w.M, w.K, lda, g.batch, sA, 16);
    const CUtensorMap tb = b_nc ? ws_make_map(B, w.N, w.K, ldb, g.batch, sB, 16) : ws_make_map(B, w.K, w.N, ldb, g.batch, sB, sh.bn);
    const int nkb = (w.K + WS_BK - 1) / WS_BK;
    DBuf ws;
    if (split > 1) {
        const int per = (nkb + split - 1) / split;
        split = (nkb + per - 1) / per;// no empty splits
        w.kb_per_split = per;
    }
    if (split > 1) {
        const long long tiles = (long long) ((w.M + sh.bm - 1) / sh.bm) * ((w.N + sh.bn - 1) / sh.bn);
        ws = DBuf(ctx, (size_t) (tiles * g.batch * split * sh.bm * sh.bn));
        w.ws = reinterpret_cast<double2 *>(ws.p);
        w.counters = ctx.ws_counters;
        w.k_splits = split;
    }
    ws_dispatch(ctx, shape, ta, tb, w, a_kc, b_nc);
    ++ctx.ws_gemms;
}

struct WsChoice {
    int shape;
    bool swap;
    int split;
    double cost;
};

// estimated SM clocks: waves x per-CTA clocks (one CTA per SM, DMMA-bound at 64 FMA/clk/SM scaled by the tile's
// fragment-load efficiency, or bound by streaming its operand bytes at the chip's ~3400 B/clk HBM rate shared by the
// CTAs of a wave), plus the split-K partial traffic
WsChoice ws_choose(const Ctx &ctx, const GemmArgs &g) {
    static const double eff[WS_SHAPE_COUNT] = {0.97, 0.88, 0.70, 0.85, 0.92, 0.68, 0.80, 0.90};
    WsChoice best{WS_64x64, false, 1, 1e300};
    const int nkb = (g.K + WS_BK - 1) / WS_BK;
    for (int swap = 0; swap < 2; ++swap) {
        const long long M = swap ? g.N : g.M, N = swap ? g.M : g.N;
        for (int t = 0; t < WS_SHAPE_COUNT; ++t) {
            const WsShapeInfo &sh = kWsShapes[t];
            if ((t == WS_16x256 || t == WS_16x192) && M > 16) continue;
            if (t == WS_48x256 && M > 48) continue;
            if (t == WS_32x256 && M > 32) continue;
            if (swap && t != WS_16x256 && t != WS_16x192 && t != WS_48x256 && t != WS_32x256) continue;// the transposed problem only to reach the skinny-M tiles
            const long long tiles = ((M + sh.bm - 1) / sh.bm) * ((N + sh.bn - 1) / sh.bn) * g.batch;
            const long long slots = ctx.sm_count;
            int cands[16], nc = 0;
            for (int split = 1; split <= 512; split *= 2) cands[nc++] = split;
            if (tiles < slots) cands[nc++] = (int) (slots / tiles), cands[nc++] = (int) ((2 * slots) / tiles);
            for (int q = 0; q < nc; ++q) {
                const int split = cands[q];
                if (split > 1 && (nkb / split < 4 || tiles > ctx.ws_counters_cap)) continue;
                const long long ctas = tiles * split;
                const double waves = (double) ((ctas + slots - 1) / slots);
                const double kb = (double) ((nkb + split - 1) / split);
                const double mma = sh.bm * sh.bn * (kb + 4.0) / (4.0 * eff[t]);                  // clocks per CTA, tensor-bound
                const double conc = (double) std::min<long long>(slots, ctas);
                // rows past the M / N edge are zero-filled by the copy engine without memory traffic
                const double rows = (double) std::min<long long>(sh.bm, M) + (double) std::min<long long>(sh.bn, N);
                const double bytes = rows * 16.0 * 8.0 * kb * conc / 3400.0 + 1500.0;            // clocks per wave, HBM-bound (+ fill latency)
                double cost = waves * (std::max(mma, bytes) + 4000.0 + 0.25 * sh.bm * sh.bn);// + per-CTA prologue / epilogue
                // split-K: every CTA writes its partial tile; the last CTA of a tile reads all of them back (~150 B/clk)
                if (split > 1) cost += 8.0 * sh.bm * sh.bn * (double) ctas / 6000.0 + 8.0 * sh.bm * sh.bn * split / 150.0 + 600.0;
                if (cost < best.cost) best = WsChoice{t, swap != 0, split, cost};
            }
        }
    }
    return best;
}

}// namespace

bool gemm_takes_tables(const Ctx &ctx, const GemmArgs &g) {
    if (ctx.shard && ctx.shard->nranks > 1) return false;// the row-sharded path assembles dense row-major results
    return ws_eligible(ctx, g);
}

void gemm(Ctx &ctx, GemmArgs g, bool a_kc, bool b_nc) {
    QCP2_REQUIRE(g.M >= 0 && g.N >= 0 && g.K >= 0, "dgemm: negative extent");
    if (g.M == 0 || g.N == 0 || g.batch == 0) return;
    // 16-byte cp.async needs every row start to be 16-byte aligned
    bool v2 = aligned16(g.A) && (g.lda % 2 == 0) && (g.strideA % 2 == 0);
    if (g.seg == nullptr) v2 = v2 && aligned16(g.B) && (g.ldb % 2 == 0) && (g.strideB % 2 == 0);
    else {
        v2 = v2 && g.seg_aligned;
        for (int s = 0; s < g.nseg; ++s) v2 = v2 && (g.seg_ldb[s] % 2 == 0);
    }
    // multi-GPU: the product is split over the ranks along its LONGER free dimension, so that the larger operand
    // is streamed only once in total.  Rows: rank r computes rows [r R, (r+1) R) of C into its slab of a W x R row
    // workspace, ONE in-place ncclAllGather assembles the product on every rank, the first M rows are copied to C.
    // Columns (N > M, beta = 0): the same on the transposed problem C^T = B^T A^T, followed by a transposing copy.
    // Every rank takes the same decision (same shapes, same threshold), so the collectives match up.
    if (ctx.shard && ctx.shard->nranks > 1 && g.seg == nullptr && g.batch == 1 && g.k_splits == 1 && g.row_off == nullptr &&
        (g.ldc == g.N || g.c_owns_rows) && 2.0 * g.M * g.N * g.K >= ctx.shard_min_flops) {
        Comm *cm = ctx.shard;
        const int W = cm->nranks;
        const bool by_cols = g.N > g.M && g.beta == 0.0 && g.N >= 2 * W;
        if (by_cols || g.M >= 2 * W) {
            const long long ext = by_cols ? g.N : g.M;           // the sharded extent
            const long long width = by_cols ? g.M : g.ldc;       // row length of the gathered workspace
            const long long R = (ext + W - 1) / W;
            const long long x0 = std::min<long long>(ext, cm->rank * R), x1 = std::min<long long>(ext, x0 + R);
            DBuf gathered(ctx, (size_t) (W * R * width));
            double *slab = gathered.p + cm->rank * R * width;
            ctx.shard = nullptr;// the slab itself is an ordinary product
            try {
                if (x1 > x0) {
                    GemmArgs part = g;
                    part.c_owns_rows = false;
                    if (!by_cols) {
                        part.M = (int) (x1 - x0);
                        part.A = g.A + (a_kc ? x0 * g.lda : x0);
                        part.C = slab;
                        if (g.beta != 0.0) {// beta term: this rank's rows of the (replicated) C
                            part.Cin = (g.Cin ? g.Cin + x0 * g.ldcin : g.C + x0 * g.ldc);
                            part.ldcin = g.Cin ? g.ldcin : g.ldc;
                        }
                        gemm(ctx, part, a_kc, b_nc);
                    } else {// slab of C^T: [x1-x0][M] = (B[:, x0:x1])^T . A^T
                        part.M = (int) (x1 - x0), part.N = g.M;
                        part.A = g.B + (b_nc ? x0 : x0 * g.ldb), part.lda = g.ldb;
                        part.B = g.A, part.ldb = g.lda;
                        part.C = slab, part.ldc = g.M;
                        part.Cin = nullptr;
                        gemm(ctx, part, /*a_kc=*/!b_nc, /*b_nc=*/!a_kc);
                    }
                }
                comm_allgather(ctx, *cm, slab, gathered.p, R * width);
                if (!by_cols) {
                    QCP2_CUDA(cudaMemcpyAsync(g.C, gathered.p, (size_t) ((long long) g.M * g.ldc) * 8, cudaMemcpyDeviceToDevice, ctx.stream));
                } else {// C[m][n] = gathered[n][m]
                    const long long oext[2] = {g.M, g.N}, istr[2] = {1, g.M};
                    if (g.ldc == g.N) {
                        permute_axpby(ctx, g.C, gathered.p, 2, oext, istr, 1.0, 0.0);
                    } else {// padded rows (c_owns_rows): transpose row by row block
                        DBuf dense(ctx, (size_t) ((long long) g.M * g.N));
                        permute_axpby(ctx, dense.p, gathered.p, 2, oext, istr, 1.0, 0.0);
                        QCP2_CUDA(cudaMemcpy2DAsync(g.C, (size_t) g.ldc * 8, dense.p, (size_t) g.N * 8, (size_t) g.N * 8, (size_t) g.M,
                                                    cudaMemcpyDeviceToDevice, ctx.stream));
                    }
                }
            } catch (...) {
                ctx.shard = cm;
                throw;
            }
            ctx.shard = cm;
            ++ctx.sharded_gemms;
            return;
        }
    }
    static const bool trace = getenv("QCP2_GEMM_TRACE") != nullptr;
    if (ws_eligible(ctx, g)) {
        WsChoice ch = ws_choose(ctx, g);
        if (g.force_tile >= 0) ch.shape = g.force_tile & 15, ch.swap = (g.force_tile & 16) != 0, ch.split = 1;
        if (g.force_split > 0) ch.split = g.force_split;
        if (trace)
            fprintf(stderr, "[qcp2 gemm] M %d N %d K %d batch %d  a_kc %d b_nc %d -> ws tile %dx%d%s split %d\n", g.M, g.N, g.K, g.batch,
                    (int) a_kc, (int) b_nc, kWsShapes[ch.shape].bm, kWsShapes[ch.shape].bn, ch.swap ? " (transposed)" : "", ch.split);
        if (ctx.log_gemms)
            ctx.gemm_log.push_back(Ctx::GemmRec{g.M, g.N, g.K, g.batch, g.beta != 0.0, kWsShapes[ch.shape].bm, kWsShapes[ch.shape].bn, ch.split});
        ws_run(ctx, g, a_kc, b_nc, ch.shape, ch.swap, ch.split);
        return;
    }
    QCP2_REQUIRE(g.row_off == nullptr, "dgemm: offset-table results need 16-byte aligned operands (warp-specialised kernel)");
    ++ctx.legacy_gemms;
    const int vec = v2 ? 2 : 1;
    if (g.seg != nullptr) {
        QCP2_REQUIRE(a_kc && b_nc, "dgemm: segmented K needs [M][K] A and [K][N] B");
        launch_dgemm_t(ctx, g, vec, true);
        return;
    }
    // tile choice: estimated SM clocks = waves x per-CTA clocks (DMMA-bound: 64 FMA/clk/SM) plus, for
    // split-K, the bandwidth-bound reduction pass (~3400 B/clk chip-wide) and its launch
    const TileChoice cand[7] = {{64, 64, 0.70}, {112, 128, 0.88}, {128, 128, 0.90}, {80, 256, 0.90}, {48, 128, 0.80},
                                {16, 256, 0.60}, {256, 16, 0.60}};
    const int nkb = (g.K + GEMM_BK - 1) / GEMM_BK;
    double best = 1e300;
    int best_t = 0, best_split = 1;
    for (int t = 0; t < 7; ++t) {
        if (t == 3 && !(a_kc && b_nc)) continue;
        if ((t == 5 && g.M > 16) || (t == 6 && g.N > 16)) continue;// skinny tiles// the 80x256 tile is instantiated for [M][K] x [K][N] only
        const long long tiles = (long long) ((g.M + cand[t].bm - 1) / cand[t].bm) * ((g.N + cand[t].bn - 1) / cand[t].bn) * g.batch;
        const int cta_per_sm = (cand[t].bm <= 64 || t == 6) ? 2 : 1;
        const long long slots = (long long) ctx.sm_count * cta_per_sm;
        // split-K candidates: powers of two and the factors that fill exactly one wave (skinny products with a
        // huge contracted dimension -- M, N <= a tile, K = o v^2 -- are bound by how many CTAs stream K)
        int cands[16], nc = 0;
        for (int split = 1; split <= 512; split *= 2) cands[nc++] = split;
        if (tiles < slots) cands[nc++] = (int) (slots / tiles), cands[nc++] = (int) std::max<long long>(1, ctx.sm_count / tiles);
        for (int q = 0; q < nc; ++q) {
            const int split = cands[q];
            if (split > 1 && (nkb / split < 4 || tiles * split > 4 * slots || t == 3)) continue;
            const long long ctas = tiles * split;
            const double waves = (double) ((ctas + slots - 1) / slots);
            // co-resident CTAs share the tensor pipe -- unless the tile is mostly padding (M or N <= 16), where the
            // product is bound by the latency of streaming K and a second CTA per SM simply doubles the bytes in flight
            const bool skinny = g.M <= 16 || g.N <= 16;
            const double conc = skinny ? 1.0 : (double) std::min<long long>(cta_per_sm, (ctas + ctx.sm_count - 1) / ctx.sm_count);
            const double kb = (double) ((nkb + split - 1) / split) + 6.0;// + prologue/epilogue overhead in k-block units
            double cost = waves * conc * cand[t].bm * cand[t].bn * kb / (4.0 * cand[t].eff);
            if (split > 1) cost += 8.0 * g.M * g.N * g.batch * (split + 1) / 3400.0 + 3000.0;
            if (cost < best) best = cost, best_t = t, best_split = split;
        }
    }
    DBuf ws;
    GemmArgs run = g;
    if (best_split > 1) {// no empty splits: with the rounded-up k-blocks per split fewer splits may already cover K
        const int per = (nkb + best_split - 1) / best_split;
        best_split = (nkb + per - 1) / per;
    }
    if (best_split > 1) {
        const long long mn = (long long) g.M * g.N;
        ws = DBuf(ctx, (size_t) (mn * best_split * g.batch));
        run.C = ws.p, run.ldc = g.N, run.strideC = mn, run.Cin = nullptr;
        run.alpha = 1.0, run.beta = 0.0;
        run.k_splits = best_split, run.kb_per_split = (nkb + best_split - 1) / best_split;
    }
    if (trace)
        fprintf(stderr, "[qcp2 gemm] M %d N %d K %d batch %d  a_kc %d b_nc %d -> tile %dx%d split %d\n", g.M, g.N, g.K, g.batch, (int) a_kc,
                (int) b_nc, cand[best_t].bm, cand[best_t].bn, best_split);
    if (ctx.log_gemms) ctx.gemm_log.push_back(Ctx::GemmRec{g.M, g.N, g.K, g.batch, g.beta != 0.0, -cand[best_t].bm, cand[best_t].bn, best_split});
    if (best_t == 6) launch_dgemm_16n(ctx, run, a_kc, b_nc, vec);
    else if (best_t == 5) launch_dgemm_16m(ctx, run, a_kc, b_nc, vec);
    else if (best_t == 4) launch_dgemm_48(ctx, run, a_kc, b_nc, vec);
    else if (best_t == 3) launch_dgemm_t(ctx, run, vec, false);
    else if (best_t == 2) launch_dgemm_128(ctx, run, a_kc, b_nc, vec);
    else if (best_t == 1) launch_dgemm_112(ctx, run, a_kc, b_nc, vec);
    else launch_dgemm_64(ctx, run, a_kc, b_nc, vec);
    if (best_split > 1) {
        const long long mn = (long long) g.M * g.N, total = mn * g.batch;
        const double *cin = g.Cin ? g.Cin : g.C;
        const long long ldcin = g.Cin ? g.ldcin : g.ldc, scin = g.Cin ? g.strideCin : g.strideC;
        splitk_reduce_kernel<<<grid_for(total, 256, ctx.sm_count), 256, 0, ctx.stream>>>(ws.p, best_split, mn, g.N, g.C, g.ldc, g.strideC,
                                                                                      cin, ldcin, scin, g.alpha, g.beta, g.batch);
        ctx.launched("splitk_reduce");
    }
}

}// namespace qcp2
