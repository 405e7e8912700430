// dgemm_ws.cuh -- general FP64 GEMM as a warp-specialised, TMA-fed DMMA kernel (sm_100a).
//
// The pipeline of t_gemm_tma.cu (one producer thread, mbarrier full/empty ring, eight consumer
// warps that never meet at a block-wide barrier) made general:
//
//  * both operands arrive through 3-D TMA tensor maps (cp.async.bulk.tensor, SASS UTMALDG) with
//    128-byte swizzle, so all four operand layouts, every M/N/K edge (hardware zero fill) and
//    strided batches are handled by the copy engine; no thread computes a global address.
//    K-contiguous operands ([M][K], [N][K]) are one box {16 k, rows} per stage; operands whose
//    contracted index is the slow one ([K][M], [K][N]) are boxes {16 cols, 16 k}, one per 16 columns.
//  * shared memory is dense (128-byte rows, no padding); bank conflicts are removed by the
//    hardware swizzle plus a permutation of k inside each 16-wide k-block: k-step s gives lane
//    t4 the index k(s,t4) = 8 (t4>>1) + (t4&1) + 2 ((t4+s)&3).  Both operands use the same
//    permutation, so the product is unchanged, and with it every 8-byte fragment load of either
//    layout touches 16 distinct 8-byte bank pairs per half warp (derivation in DESIGN.md 4.2).
//  * fragments (8-row groups) are dealt to the warps round-robin, which makes every fragment
//    address `lane offset + compile-time constant`.
//  * split-K without a second kernel or a wait: each split writes its partial tile to a
//    workspace in fragment order, bumps a per-tile counter, and the LAST CTA to arrive sums the
//    partials in split order (deterministic) and runs the epilogue.  Nobody spins.
//  * epilogue: C = alpha*acc + beta*Cin with arbitrary row/column strides (transposed results),
//    or through row/column offset tables (results whose index order interleaves the free
//    indices of A and B: the permuted targets of cc.cpp:230, 246-249) -- no re-layout pass.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <cstdint>

#include "common.cuh"

namespace qcp2 {

constexpr int WS_BK = 16;
constexpr int WS_NCW = 8;           // consumer warps
constexpr int WS_THREADS = 384;     // 2 consumer warpgroups + 1 producer warpgroup
constexpr int WS_SMEM_BUDGET = 225 * 1024;

struct WsGemmArgs {
    int M = 0, N = 0, K = 0;
    int batch = 1;
    int a_batched = 0, b_batched = 0;     // 0: the operand is shared by all batch members
    double alpha = 1.0, beta = 0.0;
    double *C = nullptr;
    const double *Cin = nullptr;           // source of the beta term (same addressing as C)
    long long c_rs = 0, c_cs = 1, strideC = 0, strideCin = 0;
    const long long *row_off = nullptr;    // optional offset tables: address = C + row_off[m] + col_off[n]
    const long long *col_off = nullptr;
    int k_splits = 1, kb_per_split = 0;
    int tiles = 0;                         // M-tiles x N-tiles; the grid is persistent: CTA b works on items b, b + gridDim.x, ...
    long long work = 0;                    // of the tiles x k_splits x batch work items (tile fastest)
    double2 *ws = nullptr;                 // split-K partials, [batch][tile][split][frag][256 threads]
    unsigned int *counters = nullptr;      // one per (batch, tile); zero before and after the launch
};

namespace wsdetail {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "WS_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra WS_DONE;\n"
            "bra WS_WAIT;\n"
            "WS_DONE:\n"
            "}\n" ::"r"(smem_u32(bar)),
            "r"(parity)
            : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                    smem_u32(dst)),
            "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

}// namespace wsdetail

// NWM x NWN consumer warps (NWM * NWN = 8), FM x FN 8x8 fragments per warp: CTA tile (NWM FM 8) x (NWN FN 8).
template<int NWM_, int NWN_, int FM_, int FN_>
struct WsShape {
    static constexpr int NWM = NWM_, NWN = NWN_, FM = FM_, FN = FN_;
    static constexpr int BM = NWM * FM * 8, BN = NWN * FN * 8;
    static constexpr int A_STAGE = BM * WS_BK, B_STAGE = BN * WS_BK;// doubles; 128-byte rows
    static constexpr int STAGE_BYTES = (A_STAGE + B_STAGE) * 8;
    static constexpr int STAGES = (WS_SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (WS_SMEM_BUDGET / STAGE_BYTES);
    static constexpr size_t SMEM = (size_t) STAGES * STAGE_BYTES + 2 * STAGES * 8 + 16;// + barriers, flag
    static_assert(NWM * NWN == WS_NCW, "eight consumer warps");
    static_assert(BM % 16 == 0 && BN % 16 == 0 && BM <= 256 && BN <= 256, "TMA boxes are 16 columns wide, at most 256 rows");
    static_assert(STAGES >= 3, "tile too large for a 3-stage ring");
    static_assert((A_STAGE * 8) % 1024 == 0 && (B_STAGE * 8) % 1024 == 0, "swizzle atoms are 1024 bytes");
};

template<class Sh, bool A_KC, bool B_NC>
__global__ void __launch_bounds__(WS_THREADS, 1)
        ws_dgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WsGemmArgs g) {
    using namespace wsdetail;
    constexpr int NWM = Sh::NWM, NWN = Sh::NWN, FM = Sh::FM, FN = Sh::FN, BM = Sh::BM, BN = Sh::BN, S = Sh::STAGES;
    constexpr int BK = WS_BK, A_STAGE = Sh::A_STAGE, B_STAGE = Sh::B_STAGE;

    // 1024-byte alignment: the swizzle pattern is a function of the shared-memory address bits 4..9 (the kernel has
    // no static shared memory, so the dynamic window starts at offset 0 of the CTA's allocation)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *As = reinterpret_cast<double *>(smem_raw);
    double *Bs = As + S * A_STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(Bs + S * B_STAGE);
    uint64_t *empty = full + S;
    unsigned int *flag = reinterpret_cast<unsigned int *>(empty + S);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_m = (g.M + BM - 1) / BM;
    const int nkb_total = (g.K + BK - 1) / BK;
    // work item w -> (tile, split, batch member); both roles walk the same sequence w = blockIdx.x, + gridDim.x, ...
    // The mbarrier ring runs on across items, so the producer is already fetching the next item's first stages while
    // the consumers are in the epilogue of the current one.

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], WS_NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= WS_NCW) {
        // ------------------------------------------------------------- producer warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;\n");
        if (warp == WS_NCW && lane == 0) {
            asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
            int stage = 0;
            uint32_t phase = 1;// fresh barriers: waiting on the preceding phase returns at once
            for (long long w = blockIdx.x; w < g.work; w += gridDim.x) {
                const int tile = (int) (w % g.tiles), split = (int) ((w / g.tiles) % g.k_splits), bz = (int) (w / ((long long) g.tiles * g.k_splits));
                const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
                const int kb0 = g.k_splits > 1 ? split * g.kb_per_split : 0;
                const int nkb = g.k_splits > 1 ? max(0, min(nkb_total, kb0 + g.kb_per_split) - kb0) : nkb_total;
                const int za = g.a_batched ? bz : 0, zb = g.b_batched ? bz : 0;
                // boxes that lie entirely past the M / N edge are not fetched: their rows of the tile only feed
                // accumulator rows / columns that are never stored
                const int a_boxes = A_KC ? 1 : min(BM / 16, (g.M - m0 + 15) / 16);
                const int b_boxes = B_NC ? min(BN / 16, (g.N - n0 + 15) / 16) : 1;
                const uint32_t bytes = (uint32_t) ((A_KC ? A_STAGE : a_boxes * 256) + (B_NC ? b_boxes * 256 : B_STAGE)) * 8u;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int k0 = (kb0 + kb) * BK;
                    mbar_wait(&empty[stage], phase);
                    mbar_expect_tx(&full[stage], bytes);
                    double *as = As + stage * A_STAGE, *bs = Bs + stage * B_STAGE;
                    if (A_KC) {
                        tma_load_3d(as, &tmA, k0, m0, za, &full[stage]);
                    } else {
                        for (int b = 0; b < a_boxes; ++b) tma_load_3d(as + b * 256, &tmA, m0 + 16 * b, k0, za, &full[stage]);
                    }
                    if (B_NC) {
                        for (int b = 0; b < b_boxes; ++b) tma_load_3d(bs + b * 256, &tmB, n0 + 16 * b, k0, zb, &full[stage]);
                    } else {
                        tma_load_3d(bs, &tmB, k0, n0, zb, &full[stage]);
                    }
                    if (++stage == S) stage = 0, phase ^= 1u;
                }
            }
        }
        return;
    }

    // ----------------------------------------------------------------- consumer warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;\n");
    const int g8 = lane >> 2, t4 = lane & 3;
    const int w_m = warp % NWM, w_n = warp / NWM;
    // fragment f of the tile = rows [8f, 8f+8); warp w_m owns the M-fragments i*NWM + w_m, warp w_n the N-fragments j*NWN + w_n

    // lane offsets (doubles) of this lane's element inside fragment 0, one per k-step (permuted k, swizzled chunk)
    int a_off[BK / 4], b_off[BK / 4];
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
        const int kk = 8 * (t4 >> 1) + (t4 & 1) + 2 * ((t4 + ks) & 3);
        // K-contiguous tile: row r = 8 f + g8 (128 bytes each), 16-byte chunk (kk>>1) ^ (r & 7)
        const int kc = g8 * 16 + ((((kk >> 1) ^ g8)) << 1) + (kk & 1);
        // 16x16 boxes: row kk, column c = 8 (f & 1) + g8 of box f >> 1, chunk (c >> 1) ^ (kk & 7); the (f & 1) term is
        // XORed in below (it flips bit 3 of the offset)
        const int mc = kk * 16 + ((((g8 >> 1) ^ (kk & 7))) << 1) + (g8 & 1);
        a_off[ks] = A_KC ? kc + w_m * 128 : (mc ^ ((NWM % 2 == 0 ? (w_m & 1) : 0) << 3)) + (NWM % 2 == 0 ? (w_m >> 1) * 256 : 0);
        b_off[ks] = !B_NC ? kc + w_n * 128 : (mc ^ ((w_n & 1) << 3)) + (w_n >> 1) * 256;
    }
    static_assert(NWN % 2 == 0, "N-fragments are dealt to an even number of warps");
    // per-fragment constant part (doubles)
    auto a_frag = [&](int i) -> int {
        if (A_KC) return i * NWM * 128;
        if (NWM % 2 == 0) return i * (NWM / 2) * 256;
        return (i >> 1) * 256;// NWM == 1: fragment i sits in box i>>1, half i&1 (XOR 8 applied at the load)
    };
    auto b_frag = [&](int j) -> int { return !B_NC ? j * NWN * 128 : j * (NWN / 2) * 256; };

    double acc[FM][FN][2];
    double fa[2][FM], fb[2][FN];
    auto ldfrag = [&](int buf, int stage, int ks) {
        const double *as = As + stage * A_STAGE, *bs = Bs + stage * B_STAGE;
#pragma unroll
        for (int i = 0; i < FM; ++i) {
            const int off = (!A_KC && NWM % 2 != 0 && (i & 1)) ? (a_off[ks] ^ 8) : a_off[ks];
            fa[buf][i] = as[off + a_frag(i)];
        }
#pragma unroll
        for (int j = 0; j < FN; ++j) fb[buf][j] = bs[b_off[ks] + b_frag(j)];
    };

    int stage = 0;
    uint32_t phase = 0;
    for (long long w = blockIdx.x; w < g.work; w += gridDim.x) {
    const int tile = (int) (w % g.tiles), split = (int) ((w / g.tiles) % g.k_splits), bz = (int) (w / ((long long) g.tiles * g.k_splits));
    const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
    const int kb0 = g.k_splits > 1 ? split * g.kb_per_split : 0;
    const int nkb = g.k_splits > 1 ? max(0, min(nkb_total, kb0 + g.kb_per_split) - kb0) : nkb_total;
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (nkb > 0) {
        mbar_wait(&full[stage], phase);
        ldfrag(0, stage, 0);
        for (int kb = 0; kb < nkb; ++kb) {
            int nstage = stage + 1;
            uint32_t nphase = phase;
            if (nstage == S) nstage = 0, nphase ^= 1u;
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                if (ks < BK / 4 - 1) {
                    ldfrag((ks + 1) & 1, stage, ks + 1);
                } else if (kb + 1 < nkb) {// first fragments of the next stage before this stage's last DMMAs
                    mbar_wait(&full[nstage], nphase);
                    ldfrag(0, nstage, 0);
                }
#pragma unroll
                for (int i = 0; i < FM; ++i)
#pragma unroll
                    for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[ks & 1][i], fb[ks & 1][j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);// every fragment of this stage has been consumed
            stage = nstage, phase = nphase;
        }
    }

    // ---- split-K: publish the partial tile; only the last CTA of this tile continues into the epilogue.
    // Fragments that lie entirely past the M / N edge (most of a skinny tile) are neither published nor reduced.
    if (g.k_splits > 1) {
        const long long slot = ((long long) bz * g.tiles + tile) * g.k_splits;
        double2 *mine = g.ws + (slot + split) * (FM * FN * 256) + tid;
        bool live[FM][FN];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) {
                live[i][j] = (m0 + (i * NWM + w_m) * 8 < g.M) && (n0 + (j * NWN + w_n) * 8 < g.N);
                if (live[i][j]) mine[(i * FN + j) * 256] = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        __threadfence();
        consumer_bar();
        if (tid == 0) {
            const unsigned int old = atomicAdd(&g.counters[(long long) bz * g.tiles + tile], 1u);
            *flag = old;
        }
        consumer_bar();
        const bool last = *flag == (unsigned int) g.k_splits - 1;
        consumer_bar();// every consumer has read the flag before a later item may overwrite it
        if (!last) continue;
        __threadfence();
        const double2 *all = g.ws + slot * (FM * FN * 256) + tid;
        constexpr long long SPLIT_STRIDE = (long long) FM * FN * 256;
        // Each element is summed in split order (the result does not depend on which CTA arrived last).  The loads are
        // issued in independent batches -- a few splits: the splits of a whole fragment row at once; many splits: sixteen
        // splits of one fragment at once -- so the walk over the partials is bandwidth- rather than latency-bound.
        if (g.k_splits <= 4) {
#pragma unroll
            for (int i = 0; i < FM; ++i) {
                double2 p[FN][4];
#pragma unroll
                for (int j = 0; j < FN; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        p[j][q] = (live[i][j] && q < g.k_splits) ? __ldcg(all + q * SPLIT_STRIDE + (i * FN + j) * 256) : make_double2(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q < g.k_splits) s0 += p[j][q].x, s1 += p[j][q].y;
                    acc[i][j][0] = s0, acc[i][j][1] = s1;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    double s0 = 0.0, s1 = 0.0;
                    if (live[i][j]) {
                        const double2 *src = all + (i * FN + j) * 256;
                        int q = 0;
                        for (; q + 16 <= g.k_splits; q += 16) {
                            double2 p[16];
#pragma unroll
                            for (int u = 0; u < 16; ++u) p[u] = __ldcg(src + (long long) (q + u) * SPLIT_STRIDE);
#pragma unroll
                            for (int u = 0; u < 16; ++u) s0 += p[u].x, s1 += p[u].y;
                        }
                        double2 p[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) p[u] = (q + u < g.k_splits) ? __ldcg(src + (long long) (q + u) * SPLIT_STRIDE) : make_double2(0.0, 0.0);
#pragma unroll
                        for (int u = 0; u < 16; ++u)
                            if (q + u < g.k_splits) s0 += p[u].x, s1 += p[u].y;
                    }
                    acc[i][j][0] = s0, acc[i][j][1] = s1;
                }
        }
        if (tid == 0) g.counters[(long long) bz * g.tiles + tile] = 0u;// ready for the next launch
    }

    // ---- epilogue: C = alpha*acc + beta*Cin
    double *C = g.C + (long long) bz * g.strideC;
    const double *Cin = g.Cin ? g.Cin + (long long) bz * g.strideCin : C;
    const bool use_beta = g.beta != 0.0;
    const bool tables = g.row_off != nullptr;
    const bool vec_ok = !tables && g.c_cs == 1 && ((g.c_rs & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(Cin) & 15) == 0);
#pragma unroll
    for (int i = 0; i < FM; ++i) {
        const int row = m0 + (i * NWM + w_m) * 8 + g8;
        if (row >= g.M) continue;
        const long long roff = tables ? g.row_off[row] : (long long) row * g.c_rs;
        // all beta-term loads of this fragment row are issued before its first store (C and Cin may alias, so the
        // compiler cannot hoist them itself): one round trip per row instead of one per 8x8 fragment
        double old0[FN], old1[FN];
        long long o0[FN], o1[FN];
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int col = n0 + (j * NWN + w_n) * 8 + 2 * t4;
            old0[j] = old1[j] = 0.0;
            o0[j] = o1[j] = 0;
            if (col >= g.N) continue;
            if (vec_ok) {
                o0[j] = roff + col;
                if (use_beta) {
                    if (col + 1 < g.N) {
                        const double2 old = *reinterpret_cast<const double2 *>(Cin + o0[j]);
                        old0[j] = old.x, old1[j] = old.y;
                    } else {
                        old0[j] = Cin[o0[j]];
                    }
                }
            } else {
                o0[j] = roff + (tables ? g.col_off[col] : (long long) col * g.c_cs);
                if (col + 1 < g.N) o1[j] = roff + (tables ? g.col_off[col + 1] : (long long) (col + 1) * g.c_cs);
                if (use_beta) {
                    old0[j] = Cin[o0[j]];
                    if (col + 1 < g.N) old1[j] = Cin[o1[j]];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int col = n0 + (j * NWN + w_n) * 8 + 2 * t4;
            if (col >= g.N) continue;
            const double v0 = g.alpha * acc[i][j][0] + g.beta * old0[j], v1 = g.alpha * acc[i][j][1] + g.beta * old1[j];
            if (vec_ok && col + 1 < g.N) {
                *reinterpret_cast<double2 *>(C + o0[j]) = make_double2(v0, v1);
            } else {
                C[o0[j]] = v0;
                if (col + 1 < g.N) C[vec_ok ? o0[j] + 1 : o1[j]] = v1;
            }
        }
    }
    }// work items
}

// ---- host side --------------------------------------------------------------------------
// 3-D tensor map {inner, outer, batch} over a row-major matrix of doubles with leading dimension ld
// and batch stride `stride` (both in doubles, both even); box {16, box_rows, 1}, 128-byte swizzle.
CUtensorMap ws_make_map(const double *base, long long inner, long long outer, long long ld, int batch, long long stride, int box_rows);
bool ws_tma_available();

template<class Sh, bool A_KC, bool B_NC>
void ws_launch_inst(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g) {
    auto kern = ws_dgemm_kernel<Sh, A_KC, B_NC>;
    static bool configured[64] = {};
    if (!configured[ctx.device & 63]) {
        QCP2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Sh::SMEM));
        configured[ctx.device & 63] = true;
    }
    const long long tiles = (long long) ((g.M + Sh::BM - 1) / Sh::BM) * ((g.N + Sh::BN - 1) / Sh::BN);
    QCP2_REQUIRE(tiles > 0 && tiles < (1LL << 31) && g.batch >= 1 && g.k_splits >= 1, "ws_dgemm: grid out of range");
    WsGemmArgs a = g;
    a.tiles = (int) tiles, a.work = tiles * g.k_splits * g.batch;
    // persistent grid: one CTA per SM (the ring takes the SM's shared memory), each walking the work items with stride gridDim.x
    const unsigned grid = (unsigned) std::min<long long>(a.work, ctx.sm_count);
    kern<<<grid, WS_THREADS, Sh::SMEM, ctx.stream>>>(tmA, tmB, a);
    ctx.launched("ws_dgemm");
}

template<class Sh>
void ws_launch_shape(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc) {
    if (a_kc && b_nc) ws_launch_inst<Sh, true, true>(ctx, tmA, tmB, g);
    else if (a_kc && !b_nc) ws_launch_inst<Sh, true, false>(ctx, tmA, tmB, g);
    else if (!a_kc && b_nc) ws_launch_inst<Sh, false, true>(ctx, tmA, tmB, g);
    else ws_launch_inst<Sh, false, false>(ctx, tmA, tmB, g);
}

// tile shapes (one translation unit each, dgemm_ws_*.cu)
enum WsShapeId { WS_80x256 = 0, WS_48x256, WS_16x256, WS_64x64, WS_112x128, WS_16x192, WS_32x256, WS_96x128, WS_SHAPE_COUNT };
using WsS80 = WsShape<2, 4, 5, 8>;  // the (T) tile: 40x64 per warp, 13 fragment loads per 40 DMMA
using WsS48 = WsShape<1, 8, 6, 4>;  // M <= 48 (packed ladder, M = o(o-1)/2 = 45): 48x32 per warp
using WsS16 = WsShape<1, 8, 2, 4>;  // M <= 16: HBM-bound streaming of the other operand
using WsS64 = WsShape<2, 4, 4, 2>;  // small products / one-wave split-K
using WsS112 = WsShape<2, 4, 7, 4>; // M = o^2 = 100
using WsS32 = WsShape<1, 8, 4, 4>;  // 16 < M <= 32 (a spin class of i<j pairs: M = 25 at o = 10)
using WsS96 = WsShape<2, 4, 6, 4>;  // 80 < M <= 96 (AO->MO at n = 86: 90 % of the rows are real, 77 % on the 112-row tile)
using WsS16N = WsShape<1, 8, 2, 3>; // M <= 16, N = v = 162 and the like: 16x192 wastes less of the tensor pipe than 16x256
struct WsShapeInfo {
    int bm, bn, stages;
};
extern const WsShapeInfo kWsShapes[WS_SHAPE_COUNT];
void ws_launch_80x256(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc);
void ws_launch_48x256(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc);
void ws_launch_16x256(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc);
void ws_launch_64x64(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc);
void ws_launch_112x128(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc);
void ws_launch_16x192(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc);
void ws_launch_32x256(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc);
void ws_launch_96x128(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc);

}// namespace qcp2
