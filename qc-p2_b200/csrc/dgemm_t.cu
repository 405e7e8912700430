// dgemm_t.cu -- the 80x256 CTA tile used by the (T) triples build: M = v = 400 is covered by
// exactly five 80-row tiles, K-contiguous A (built per triple), N-contiguous B rows that are
// either one matrix or the six-segment concatenation described in dgemm.cuh.
#include <cstdlib>

#include "dgemm.cuh"

namespace qcp2 {

// QCP2_T_VARIANT (experiments only): 0 = copies issued right after the barrier, 1 = after the
// first k-step (default), 2 = 16 warps with 40x32 warp tiles.
static int variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QCP2_T_VARIANT");
        v = e ? atoi(e) : 1;
    }
    return v;
}

template<bool SEG>
static void pick(Ctx &ctx, const GemmArgs &g, int vec) {
    const int var = variant();
    if (vec == 2) {
        if (var == 0) launch_dgemm_inst<TileT, true, true, 2, SEG, 0>(ctx, g);
        else if (var == 2) launch_dgemm_inst<TileT16, true, true, 2, SEG, 1>(ctx, g);
        else launch_dgemm_inst<TileT, true, true, 2, SEG, 1>(ctx, g);
    } else {
        launch_dgemm_inst<TileT, true, true, 1, SEG, 1>(ctx, g);
    }
}

void launch_dgemm_t(Ctx &ctx, const GemmArgs &g, int vec, bool seg) {
    if (seg) pick<true>(ctx, g, vec);
    else pick<false>(ctx, g, vec);
}

}// namespace qcp2
