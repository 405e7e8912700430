// t_gemm_tma.cu -- the (T) triples-build GEMM as a warp-specialised, TMA-fed DMMA kernel.
//
// X[t][a][p] = sum_k A[t][a][k] * B_t[k][p]        M = v, N = v(v-1)/2 (b<c packed), K = 3(v+o)
//
// One producer thread streams the operands with bulk-tensor copies (cp.async.bulk, SASS UBLKCP)
// into a 4-stage shared-memory ring guarded by mbarriers; eight consumer warps (warp tile
// 40x64, CTA tile 80x256) turn them into DMMA.8x8x4 without ever meeting at a block-wide
// barrier, so the two warps of a scheduler drift apart and cover each other's stage-boundary
// bubbles.  setmaxnreg moves the producer warpgroup's registers to the consumers.
//
//  * A is pre-tiled by build_A (triples.cu) as [m-tile][k-block][80][20]: exactly the padded
//    shared-memory image of a stage, so one 12.8 KB bulk copy fills it.
//  * B rows come straight from the packed tensors through the six-segment descriptor
//    (dgemm.cuh); each k-row of a stage is one bulk copy of up to 2 KB into a 260-double
//    (bank-conflict-free) row.  Rows past a segment's end are not copied: the slot then holds
//    zeros or finite stale data, which the matching zero columns of A annihilate.
#include <cstdint>

#include "dgemm.cuh"
#include "triples.cuh"

namespace qcp2 {

namespace {

constexpr int BM = TG_BM, BN = TG_BN, BK = GEMM_BK, S = TG_STAGES;
constexpr int ALD = TG_ALD, BLD = BN + 4;
constexpr int A_STAGE = BM * ALD, B_STAGE = BK * BLD;// doubles
constexpr int A_BYTES = A_STAGE * 8;
constexpr int NCW = 8;         // consumer warps
constexpr int THREADS = 384;   // 2 consumer warpgroups + 1 producer warpgroup
constexpr int WM = 40, WN = 64, FM = WM / 8, FN = WN / 8;

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
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra DONE;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}\n" ::"r"(smem_u32(bar)),
            "r"(parity)
            : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) t_gemm_tma_kernel(const TmaGemmArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *As = reinterpret_cast<double *>(smem_raw);
    double *Bs = As + S * A_STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(Bs + S * B_STAGE);
    uint64_t *empty = full + S;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mtiles = (g.M + BM - 1) / BM;
    const int mt = blockIdx.x % mtiles, n0 = (blockIdx.x / mtiles) * BN;
    const long long bz = blockIdx.z;
    const int nkb = g.nkb;

    {// pad rows / tail columns must hold finite values
        double2 *z = reinterpret_cast<double2 *>(Bs);
        for (int q = tid; q < S * B_STAGE / 2; q += THREADS) z[q] = make_double2(0.0, 0.0);
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");// generic zero-fill before async-proxy writes
    __syncthreads();

    if (warp >= NCW) {
        // ------------------------------------------------------------- producer warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;\n");
        if (warp == NCW && lane == 0) {
            const double *Asrc = g.At + bz * g.strideA + (long long) mt * nkb * A_STAGE;
            const SegDesc sd = g.seg[bz];
            const int ncols = min(BN, g.N - n0);
            const uint32_t rowbytes = (uint32_t) ncols * 8u;
            int stage = 0;
            uint32_t phase = 1;// fresh barriers: waiting on the preceding phase returns at once
            for (int kb = 0; kb < nkb; ++kb) {
                int s = 0;
#pragma unroll
                for (int q = 1; q < GEMM_MAX_SEG; ++q)
                    if (q < g.nseg && kb >= g.seg_kb_begin[q]) s = q;
                const int lk0 = (kb - g.seg_kb_begin[s]) * BK;
                const int rows = max(0, min(BK, g.seg_rows[s] - lk0));
                mbar_wait(&empty[stage], phase);
                mbar_expect_tx(&full[stage], (uint32_t) A_BYTES + (uint32_t) rows * rowbytes);
                bulk_g2s(As + stage * A_STAGE, Asrc + (long long) kb * A_STAGE, A_BYTES, &full[stage]);
                double *bdst = Bs + stage * B_STAGE;
                if (s == 3 && g.merged_sub_rows > 0) {// merged hole block: row lk0 + r lives in sub-segment (lk0 + r) / o
                    for (int r = 0; r < rows; ++r) {
                        const int lr = lk0 + r, sub = lr / g.merged_sub_rows, loc = lr - sub * g.merged_sub_rows;
                        bulk_g2s(bdst + r * BLD, sd.B[3 + sub] + (long long) loc * g.seg_ldb[3] + n0, rowbytes, &full[stage]);
                    }
                } else {
                    const double *brow = sd.B[s] + (long long) lk0 * g.seg_ldb[s] + n0;
                    for (int r = 0; r < rows; ++r) bulk_g2s(bdst + r * BLD, brow + (long long) r * g.seg_ldb[s], rowbytes, &full[stage]);
                }
                if (++stage == S) stage = 0, phase ^= 1u;
            }
        }
        return;
    }

    // ----------------------------------------------------------------- consumer warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;\n");
    const int g8 = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp & 1) * WM, wn0 = (warp >> 1) * WN;
    const double *a_lane = As + (wm0 + g8) * ALD + t4;
    const double *b_lane = Bs + t4 * BLD + wn0 + g8;

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double fa[2][FM], fb[2][FN];
    auto ldfrag = [&](int buf, int stage, int ks) {
        const double *as = a_lane + stage * A_STAGE + ks * 4;
        const double *bs = b_lane + stage * B_STAGE + ks * 4 * BLD;
#pragma unroll
        for (int i = 0; i < FM; ++i) fa[buf][i] = as[i * 8 * ALD];
#pragma unroll
        for (int j = 0; j < FN; ++j) fb[buf][j] = bs[j * 8];
    };

    int stage = 0;
    uint32_t phase = 0;
    mbar_wait(&full[0], 0);
    ldfrag(0, 0, 0);
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
                for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[ks & 1][i], fb[ks & 1][j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);// every fragment of this stage has been consumed
        stage = nstage, phase = nphase;
    }

    double *C = g.C + bz * g.strideC;
#pragma unroll
    for (int i = 0; i < FM; ++i) {
        const int row = mt * BM + wm0 + i * 8 + g8;
        if (row >= g.M) continue;
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int col = n0 + wn0 + j * 8 + 2 * t4;
            if (col >= g.N) continue;// N is even, so col+1 < N as well
            *reinterpret_cast<double2 *>(C + (long long) row * g.ldc + col) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

}// namespace

size_t t_gemm_tma_smem_bytes() { return (size_t) S * (A_STAGE + B_STAGE) * 8 + 2 * S * 8; }

void launch_t_gemm_tma(Ctx &ctx, const TmaGemmArgs &g, int batch) {
    QCP2_REQUIRE(g.N % 2 == 0 && g.ldc % 2 == 0 && g.strideC % 2 == 0, "t_gemm_tma: N, ldc, strideC must be even");
    static bool configured[64] = {};
    const size_t smem = t_gemm_tma_smem_bytes();
    if (!configured[ctx.device & 63]) {
        QCP2_CUDA(cudaFuncSetAttribute(t_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        configured[ctx.device & 63] = true;
    }
    const long long tiles = (long long) ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
    QCP2_REQUIRE(tiles > 0 && tiles < (1LL << 31) && batch >= 1 && batch <= 65535, "t_gemm_tma: grid out of range");
    dim3 grid((unsigned) tiles, 1, (unsigned) batch);
    t_gemm_tma_kernel<<<grid, THREADS, smem, ctx.stream>>>(g);
    ctx.launched("t_gemm_tma");
}

}// namespace qcp2
