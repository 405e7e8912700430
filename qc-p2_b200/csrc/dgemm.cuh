// dgemm.cuh -- FP64 tensor-core GEMM for sm_100a.
//
// The only FP64 tensor instruction on sm_100a is DMMA.8x8x4 (mma.sync.m8n8k4.f64; tcgen05
// has no f64 kind), so this is a register-fragment kernel: global -> shared memory with a
// multi-stage cp.async pipeline, 8-byte fragment loads from padded (bank-conflict-free)
// shared tiles, one DMMA per 8x8x4 block, FP64 accumulators in registers.
//
// Row-major operands.  A is either [M][K] (A_KC: K contiguous) or [K][M]; B is either
// [K][N] (B_NC: N contiguous) or [N][K]; C is [M][N].  All edges are predicated with
// zero-fill, nothing is assumed about M, N, K or the leading dimensions except that VEC=2
// kernels need 16-byte aligned rows (the launcher checks and falls back to VEC=1).
//
// SEG mode (used by the (T) triples build): the K dimension is the concatenation of up to
// 6 row blocks of B living in different tensors (vovv[:,i], vovv[:,j], vovv[:,k], T2[i],
// T2[j], T2[k]); each block is padded to a multiple of BK in the K index space and its row
// pointers are taken from a per-batch descriptor, so nothing is ever copied together.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace qcp2 {

constexpr int GEMM_BK = 16;
constexpr int GEMM_MAX_SEG = 6;

struct SegDesc {
    const double *B[GEMM_MAX_SEG];
};

struct GemmArgs {
    const double *A = nullptr;
    long long lda = 0, strideA = 0;
    const double *B = nullptr;
    long long ldb = 0, strideB = 0;
    double *C = nullptr;
    long long ldc = 0, strideC = 0;
    const double *Cin = nullptr;// source of the beta term (defaults to C)
    long long ldcin = 0, strideCin = 0;
    int M = 0, N = 0, K = 0;
    double alpha = 1.0, beta = 0.0;
    int batch = 1;
    // split-K: grid.y = k_splits, split s handles k-blocks [s*kb_per_split, ...) and writes its raw
    // partial product to C + (batch*k_splits + s) * strideC (the launcher points C at a workspace)
    int k_splits = 1, kb_per_split = 0;
    // the caller owns every element of the M x ldc region behind C (padding columns included), so whole rows may
    // be exchanged between ranks by the row-sharded path
    bool c_owns_rows = false;
    // optional result addressing through offset tables (device): &C[m][n] = C + row_off[m] + col_off[n]; the same
    // addressing applies to Cin.  Only the warp-specialised kernel (dgemm_ws.cuh) implements it.
    const long long *row_off = nullptr, *col_off = nullptr;
    // tests / tuning: force a warp-specialised tile (WsShapeId, +8 = transposed problem) and a split-K factor
    int force_tile = -1, force_split = 0;
    // SEG mode
    const SegDesc *seg = nullptr;
    int nseg = 0;
    bool seg_aligned = true;// every segment base pointer is 16-byte aligned
    int seg_kb_begin[GEMM_MAX_SEG + 1] = {0, 0, 0, 0, 0, 0, 0};
    int seg_rows[GEMM_MAX_SEG] = {0, 0, 0, 0, 0, 0};
    long long seg_ldb[GEMM_MAX_SEG] = {0, 0, 0, 0, 0, 0};
};

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc, int src_bytes) {
    unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, int src_bytes) {
    unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template<int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Copy a ROWS x COLS tile (COLS contiguous in global memory) into padded shared memory,
// zero-filling everything outside [rows_lim) x [cols_lim).
template<int ROWS, int COLS, int LD, int VEC, int NT>
__device__ __forceinline__ void load_tile(double *s, const double *gbase, long long ld, int row0, int col0,
                                          int rows_lim, int cols_lim, int tid) {
    constexpr int CPR = COLS / VEC;
    constexpr int TOTAL = ROWS * CPR;
#pragma unroll
    for (int it = 0; it < (TOTAL + NT - 1) / NT; ++it) {
        const int ch = it * NT + tid;
        if ((TOTAL % NT != 0) && ch >= TOTAL) break;
        const int r = ch / CPR, c = (ch % CPR) * VEC;
        const int gr = row0 + r, gc = col0 + c;
        int nbytes = 0;
        const double *src = gbase;
        if (gr < rows_lim && gc < cols_lim) {
            nbytes = (VEC == 2 && gc + 1 < cols_lim) ? 16 : 8;
            src = gbase + (long long) gr * ld + gc;
        }
        if (VEC == 2) cp_async16(s + r * LD + c, src, nbytes);
        else cp_async8(s + r * LD + c, src, nbytes);
    }
}

template<int BM, int BN, int WM, int WN, int STAGES>
struct GemmTile {
    static constexpr int kBM = BM, kBN = BN, kWM = WM, kWN = WN, kStages = STAGES;
    static constexpr int kThreads = (BM / WM) * (BN / WN) * 32;
    template<bool A_KC, bool B_NC>
    static constexpr size_t smem_bytes() {
        return (size_t) STAGES * ((A_KC ? BM * (GEMM_BK + 4) : GEMM_BK * (BM + 4)) +
                                  (B_NC ? GEMM_BK * (BN + 4) : BN * (GEMM_BK + 4))) *
               sizeof(double);
    }
};

template<class Tile, bool A_KC, bool B_NC, int VEC, bool SEG, int SCHED = 1>
__global__ void __launch_bounds__(Tile::kThreads, 1) dgemm_kernel(const GemmArgs g) {
    constexpr int BM = Tile::kBM, BN = Tile::kBN, WM = Tile::kWM, WN = Tile::kWN, STAGES = Tile::kStages;
    constexpr int BK = GEMM_BK;
    constexpr int NWM = BM / WM, NT = Tile::kThreads;
    constexpr int A_LD = A_KC ? BK + 4 : BM + 4, A_ROWS = A_KC ? BM : BK;
    constexpr int B_LD = B_NC ? BN + 4 : BK + 4, B_ROWS = B_NC ? BK : BN;
    constexpr int A_STAGE = A_ROWS * A_LD, B_STAGE = B_ROWS * B_LD;
    constexpr int FM = WM / 8, FN = WN / 8;
    static_assert(!SEG || B_NC, "segmented K needs N-contiguous B rows");
    static_assert(BM % 16 == 0 && BN % 16 == 0 && WM % 8 == 0 && WN % 8 == 0, "tile shape");

    extern __shared__ __align__(16) double smem[];
    double *As = smem, *Bs = smem + STAGES * A_STAGE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp % NWM) * WM, wn0 = (warp / NWM) * WN;

    const int tiles_m = (g.M + BM - 1) / BM;
    const int m0 = (blockIdx.x % tiles_m) * BM, n0 = (blockIdx.x / tiles_m) * BN;
    const long long bz = blockIdx.z;
    const double *A = g.A + bz * g.strideA;
    const double *B = SEG ? nullptr : g.B + bz * g.strideB;
    const int nkb_total = (g.K + BK - 1) / BK;
    const int kb0 = g.k_splits > 1 ? (int) blockIdx.y * g.kb_per_split : 0;
    const int nkb = g.k_splits > 1 ? max(0, min(nkb_total, kb0 + g.kb_per_split) - kb0) : nkb_total;

    auto load_stage = [&](int kb, int stage) {
        double *as = As + stage * A_STAGE, *bs = Bs + stage * B_STAGE;
        const int k0 = (kb0 + kb) * BK;
        if (A_KC) load_tile<BM, BK, A_LD, VEC, NT>(as, A, g.lda, m0, k0, g.M, g.K, tid);
        else load_tile<BK, BM, A_LD, VEC, NT>(as, A, g.lda, k0, m0, g.K, g.M, tid);
        if (SEG) {
            int s = 0;
#pragma unroll
            for (int q = 1; q < GEMM_MAX_SEG; ++q)
                if (q < g.nseg && kb0 + kb >= g.seg_kb_begin[q]) s = q;
            const int lk0 = (kb0 + kb - g.seg_kb_begin[s]) * BK;
            load_tile<BK, BN, B_LD, VEC, NT>(bs, g.seg[bz].B[s], g.seg_ldb[s], lk0, n0, g.seg_rows[s], g.N, tid);
        } else if (B_NC) {
            load_tile<BK, BN, B_LD, VEC, NT>(bs, B, g.ldb, k0, n0, g.K, g.N, tid);
        } else {
            load_tile<BN, BK, B_LD, VEC, NT>(bs, B, g.ldb, n0, k0, g.N, g.K, tid);
        }
    };

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }

    // SCHED 0: issue the next stage's copies right after the barrier; SCHED 1: after the first
    // k-step's DMMAs, so the tensor pipe restarts as soon as the first fragments are loaded
    // and the copy address arithmetic hides behind DMMA issue.
    for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kb + STAGES - 1;
        if (SCHED == 0) {
            if (nk < nkb) load_stage(nk, nk % STAGES);
            cp_async_commit();
        }
        const double *as = As + (kb % STAGES) * A_STAGE;
        const double *bs = Bs + (kb % STAGES) * B_STAGE;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double a[FM], b[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i)
                a[i] = A_KC ? as[(wm0 + i * 8 + g8) * A_LD + ks * 4 + t4] : as[(ks * 4 + t4) * A_LD + wm0 + i * 8 + g8];
#pragma unroll
            for (int j = 0; j < FN; ++j)
                b[j] = B_NC ? bs[(ks * 4 + t4) * B_LD + wn0 + j * 8 + g8] : bs[(wn0 + j * 8 + g8) * B_LD + ks * 4 + t4];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            if (SCHED == 1 && ks == 0) {
                if (nk < nkb) load_stage(nk, nk % STAGES);
                cp_async_commit();
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: C = alpha*acc + beta*Cin
    double *C = g.C + (g.k_splits > 1 ? (bz * g.k_splits + blockIdx.y) : bz) * g.strideC;
    const double *Cin = g.Cin ? g.Cin + bz * g.strideCin : C;
    const long long ldcin = g.Cin ? g.ldcin : g.ldc;
    const bool vec_ok = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && ((ldcin & 1) == 0) &&
                        ((reinterpret_cast<uintptr_t>(Cin) & 15) == 0);
    const bool use_beta = g.beta != 0.0;
#pragma unroll
    for (int i = 0; i < FM; ++i) {
        const int row = m0 + wm0 + i * 8 + g8;
        if (row >= g.M) continue;
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int col = n0 + wn0 + j * 8 + 2 * t4;
            if (col >= g.N) continue;
            double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
            double *dst = C + (long long) row * g.ldc + col;
            const double *src = Cin + (long long) row * ldcin + col;
            if (col + 1 < g.N && vec_ok) {
                if (use_beta) {
                    const double2 old = *reinterpret_cast<const double2 *>(src);
                    v0 += g.beta * old.x, v1 += g.beta * old.y;
                }
                *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
            } else {
                if (use_beta) v0 += g.beta * src[0];
                dst[0] = v0;
                if (col + 1 < g.N) {
                    if (use_beta) v1 += g.beta * src[1];
                    dst[1] = v1;
                }
            }
        }
    }
}

// host-side launcher for one instantiation
template<class Tile, bool A_KC, bool B_NC, int VEC, bool SEG, int SCHED = 1>
void launch_dgemm_inst(Ctx &ctx, const GemmArgs &g) {
    auto kern = dgemm_kernel<Tile, A_KC, B_NC, VEC, SEG, SCHED>;
    constexpr size_t smem = Tile::template smem_bytes<A_KC, B_NC>();
    static bool configured[64] = {};
    if (!configured[ctx.device & 63]) {
        QCP2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        configured[ctx.device & 63] = true;
    }
    const long long tiles = (long long) ((g.M + Tile::kBM - 1) / Tile::kBM) * ((g.N + Tile::kBN - 1) / Tile::kBN);
    QCP2_REQUIRE(tiles > 0 && tiles < (1LL << 31) && g.batch >= 1 && g.batch <= 65535, "dgemm: grid out of range");
    dim3 grid((unsigned) tiles, (unsigned) std::max(g.k_splits, 1), (unsigned) g.batch);
    kern<<<grid, Tile::kThreads, smem, ctx.stream>>>(g);
    ctx.launched("dgemm");
}

using Tile64 = GemmTile<64, 64, 32, 32, 3>;
using Tile48 = GemmTile<48, 128, 48, 16, 3>;
// skinny products (one free dimension = o <= 16, the other o v^2): HBM-bound streaming of the big operand
using Tile16M = GemmTile<16, 256, 16, 32, 3>;
using Tile16N = GemmTile<256, 16, 32, 16, 3>;// M = o(o-1)/2 = 45 rows of the packed particle-particle ladder
using Tile128 = GemmTile<128, 128, 64, 32, 3>;
using Tile112 = GemmTile<112, 128, 56, 32, 3>;// skinny M = o^2 = 100 (particle-particle ladder)
using TileT = GemmTile<80, 256, 40, 64, 3>;
using TileT16 = GemmTile<80, 256, 40, 32, 3>;// 16-warp experiment

// implemented in dgemm_64.cu / dgemm_128.cu / dgemm_t.cu (one translation unit per tile so
// the instantiations compile in parallel)
void launch_dgemm_64(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec);
void launch_dgemm_48(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec);
void launch_dgemm_16m(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec);
void launch_dgemm_16n(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec);
void launch_dgemm_128(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec);
void launch_dgemm_112(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec);
void launch_dgemm_t(Ctx &ctx, const GemmArgs &g, int vec, bool seg);

// dgemm.cu: alignment analysis + tile choice
void gemm(Ctx &ctx, GemmArgs g, bool a_kc, bool b_nc);
// true if gemm() will run g on the warp-specialised TMA kernel (the only one with offset-table results)
bool gemm_takes_tables(const Ctx &ctx, const GemmArgs &g);

}// namespace qcp2
