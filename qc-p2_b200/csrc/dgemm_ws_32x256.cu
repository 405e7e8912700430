// dgemm_ws_32x256.cu -- the four operand layouts of the warp-specialised TMA/DMMA GEMM for the 32x256 CTA tile.
#include "dgemm_ws.cuh"

namespace qcp2 {

void ws_launch_32x256(Ctx &ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const WsGemmArgs &g, bool a_kc, bool b_nc) {
    ws_launch_shape<WsS32>(ctx, tmA, tmB, g, a_kc, b_nc);
}

}// namespace qcp2
