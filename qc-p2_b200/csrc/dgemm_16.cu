// dgemm_16.cu -- instantiations of the DMMA GEMM for the skinny tiles 16x256 (M <= 16) and 256x16 (N <= 16).
#include "dgemm.cuh"

namespace qcp2 {

template<bool A_KC, bool B_NC>
static void pick_vec_m(Ctx &ctx, const GemmArgs &g, int vec) {
    if (vec == 2) launch_dgemm_inst<Tile16M, A_KC, B_NC, 2, false>(ctx, g);
    else launch_dgemm_inst<Tile16M, A_KC, B_NC, 1, false>(ctx, g);
}

void launch_dgemm_16m(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec) {
    if (a_kc && b_nc) pick_vec_m<true, true>(ctx, g, vec);
    else if (a_kc && !b_nc) pick_vec_m<true, false>(ctx, g, vec);
    else if (!a_kc && b_nc) pick_vec_m<false, true>(ctx, g, vec);
    else pick_vec_m<false, false>(ctx, g, vec);
}

}// namespace qcp2

namespace qcp2 {

template<bool A_KC, bool B_NC>
static void pick_vec_n(Ctx &ctx, const GemmArgs &g, int vec) {
    if (vec == 2) launch_dgemm_inst<Tile16N, A_KC, B_NC, 2, false>(ctx, g);
    else launch_dgemm_inst<Tile16N, A_KC, B_NC, 1, false>(ctx, g);
}

void launch_dgemm_16n(Ctx &ctx, const GemmArgs &g, bool a_kc, bool b_nc, int vec) {
    if (a_kc && b_nc) pick_vec_n<true, true>(ctx, g, vec);
    else if (a_kc && !b_nc) pick_vec_n<true, false>(ctx, g, vec);
    else if (!a_kc && b_nc) pick_vec_n<false, true>(ctx, g, vec);
    else pick_vec_n<false, false>(ctx, g, vec);
}

}// namespace qcp2

