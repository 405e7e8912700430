// gpu.h -- process-wide qcp2 context for the host mirror (one GPU, host pointer mode).
#pragma once
#include <stdexcept>
#include <string>

#include "../../include/qcp2_b200.h"

inline qcp2_ctx *&gpu_slot() {
    static qcp2_ctx *c = nullptr;
    return c;
}
inline qcp2_ctx *gpu() {
    qcp2_ctx *&c = gpu_slot();
    if (!c) {
        const int rc = qcp2_create(0, &c);
        if (rc != 0) throw std::runtime_error("qcp2_b200: no usable B200 device (status " + std::to_string(rc) + "); there is no CPU fallback");
    }
    return c;
}
inline void gpu_check(int rc) {
    if (rc != 0) throw std::runtime_error(std::string("qcp2_b200: ") + qcp2_last_error(gpu()));
}
inline void gpu_release() {
    qcp2_ctx *&c = gpu_slot();
    if (c) qcp2_destroy(c), c = nullptr;
}
