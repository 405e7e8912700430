// dgemm_ws.cu -- host side of the warp-specialised TMA/DMMA GEMM: tensor-map encoding through the
// driver entry point (no link-time dependency on libcuda) and the shape table.
#include <cudaTypedefs.h>

#include <mutex>

#include "dgemm_ws.cuh"

namespace qcp2 {

const WsShapeInfo kWsShapes[WS_SHAPE_COUNT] = {{WsS80::BM, WsS80::BN, WsS80::STAGES},
                                               {WsS48::BM, WsS48::BN, WsS48::STAGES},
                                               {WsS16::BM, WsS16::BN, WsS16::STAGES},
                                               {WsS64::BM, WsS64::BN, WsS64::STAGES},
                                               {WsS112::BM, WsS112::BN, WsS112::STAGES},
                                               {WsS16N::BM, WsS16N::BN, WsS16N::STAGES},
                                               {WsS32::BM, WsS32::BN, WsS32::STAGES},
                                               {WsS96::BM, WsS96::BN, WsS96::STAGES}};

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}
}// namespace

bool ws_tma_available() { return encode_fn() != nullptr; }

CUtensorMap ws_make_map(const double *base, long long inner, long long outer, long long ld, int batch, long long stride, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    QCP2_REQUIRE(enc != nullptr, "ws_dgemm: cuTensorMapEncodeTiled is not available from this driver");
    QCP2_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 1) == 0 && (stride & 1) == 0 && ld >= inner,
                 "ws_dgemm: operand is not 16-byte aligned");
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t) inner, (cuuint64_t) outer, (cuuint64_t) (batch > 1 && stride > 0 ? batch : 1)};
    // a unit batch dimension still needs a legal stride
    const long long zstride = (batch > 1 && stride > 0) ? stride : std::max<long long>(ld * outer, 2);
    const cuuint64_t strides[2] = {(cuuint64_t) ld * 8, (cuuint64_t) zstride * 8};
    const cuuint32_t box[3] = {16, (cuuint32_t) box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
        throw Error("ws_dgemm: cuTensorMapEncodeTiled failed with code " + std::to_string((int) rc) + " (inner " + std::to_string(inner) +
                    ", outer " + std::to_string(outer) + ", ld " + std::to_string(ld) + ", batch " + std::to_string(batch) + ", stride " +
                    std::to_string(stride) + ")");
    return m;
}

}// namespace qcp2
