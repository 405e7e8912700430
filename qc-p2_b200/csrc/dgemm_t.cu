// dgemm_t.cu -- the 80x256 CTA tile used by the (T) triples build: M = v = 400 is covered by
// exactly five 80-row tiles, K-contiguous A (built per triple), N-contiguous B rows that are
// either one matrix or the six-segment concatenation described in dgemm.cuh.
#include "dgemm.cuh"

namespace qcp2 {

void launch_dgemm_t(Ctx &ctx, const GemmArgs &g, int vec, bool seg) {
    if (seg) {
        if (vec == 2) launch_dgemm_inst<TileT, true, true, 2, true>(ctx, g);
        else launch_dgemm_inst<TileT, true, true, 1, true>(ctx, g);
    } else {
        if (vec == 2) launch_dgemm_inst<TileT, true, true, 2, false>(ctx, g);
        else launch_dgemm_inst<TileT, true, true, 1, false>(ctx, g);
    }
}

}// namespace qcp2
