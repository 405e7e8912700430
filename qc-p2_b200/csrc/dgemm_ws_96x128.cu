// dgemm_ws_96x128.cu -- the four operand layouts of the warp-specialised TMA/DMMA GEMM for the 96x128 CTA tile.
#include "dgemm_ws.cuh"

namespace qcp2 {

void ws_launch_96x128(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc) {
    ws_launch_shape<WsS96>(ctx, tmA, tmB, g, a_kc, b_nc);
}

}// namespace qcp2
