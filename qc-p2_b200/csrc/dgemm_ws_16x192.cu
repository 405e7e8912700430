// dgemm_ws_16x192.cu -- the four operand layouts of the warp-specialised TMA/DMMA GEMM for the 16x192 CTA tile.
#include "dgemm_ws.cuh"

namespace qcp2 {

void ws_launch_16x192(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc) {
    ws_launch_shape<WsS16N>(ctx, tmA, tmB, g, a_kc, b_nc);
}

}// namespace qcp2
