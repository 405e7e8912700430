// dgemm.cu -- alignment analysis and tile selection for the DMMA GEMM.
#include "dgemm.cuh"

namespace qcp2 {

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

void gemm(Ctx &ctx, GemmArgs g, bool a_kc, bool b_nc) {
    QCP2_REQUIRE(g.M >= 0 && g.N >= 0 && g.K >= 0, "dgemm: negative extent");
    if (g.M == 0 || g.N == 0 || g.batch == 0) return;
    // 16-byte cp.async needs every row start to be 16-byte aligned
    bool v2 = aligned16(g.A) && (g.lda % 2 == 0) && (g.strideA % 2 == 0);
    if (g.seg == nullptr) v2 = v2 && aligned16(g.B) && (g.ldb % 2 == 0) && (g.strideB % 2 == 0);
    else
    {
        v2 = v2 && g.seg_aligned;
        for (int s = 0; s < g.nseg; ++s) v2 = v2 && (g.seg_ldb[s] % 2 == 0);
    }
    const int vec = v2 ? 2 : 1;
    if (g.seg != nullptr) {
        QCP2_REQUIRE(a_kc && b_nc, "dgemm: segmented K needs [M][K] A and [K][N] B");
        launch_dgemm_t(ctx, g, vec, true);
        return;
    }
    const double work = (double) g.M * (double) g.N * (double) g.batch;
    // the 128x128 tile pays off once it can fill the machine about twice over
    const bool big = work >= 2.0 * 128.0 * 128.0 * ctx.sm_count && g.M > 64 && g.N > 64;
    if (big && a_kc && b_nc && g.M % 80 == 0 && g.M % 128 != 0 && g.N >= 1024) launch_dgemm_t(ctx, g, vec, false);
    else if (big) launch_dgemm_128(ctx, g, a_kc, b_nc, vec);
    else launch_dgemm_64(ctx, g, a_kc, b_nc, vec);
}

}// namespace qcp2
