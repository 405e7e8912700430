// dgemm_48.cu -- instantiations of the DMMA GEMM for the 48x128 CTA tile (all four operand layouts).
#include "dgemm.cuh"

namespace qcp2 {

template<bool A_KC, bool B_NC>
static void pick_vec(Ctx &ctx, const GemmArgs &g, int vec) {
    if (vec == 2) launch_dgemm_inst<Tile48, A_KC, B_NC, 2, false>(ctx, g);
    else launch_dgemm_inst<Tile48, A_KC, B_NC, 1, false>(ctx, g);
}

void launch_dgemm_48(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec) {
    if (a_kc && b_nc) pick_vec<true, true>(ctx, g, vec);
    else if (a_kc && !b_nc) pick_vec<true, false>(ctx, g, vec);
    else if (!a_kc && b_nc) pick_vec<false, true>(ctx, g, vec);
    else pick_vec<false, false>(ctx, g, vec);
}

}// namespace qcp2
