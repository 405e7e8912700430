// common.cuh -- context, error handling and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace qcp2 {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define QCP2_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::qcp2::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                                __FILE__ + ":" + std::to_string(__LINE__));                         \
    } while (0)

#define QCP2_REQUIRE(cond, msg)                       \
    do {                                              \
        if (!(cond)) throw ::qcp2::Error(std::string(msg)); \
    } while (0)

// Re-layout cache (contract.cu): while it is installed, operand re-layouts of the tensors registered
// as constant (the integral blocks inside the CCSD loop) are computed once and reused.
struct PermCache {
    struct Entry {
        const double *src;
        std::string sig;
        double *buf;
    };
    std::vector<const double *> constants;
    std::vector<Entry> entries;
    bool frozen = false;// no new entries (while the iteration is being captured into a CUDA graph)
    bool is_constant(const double *p) const {
        for (const double *c : constants)
            if (c == p) return true;
        return false;
    }
};

// Execution context: one device, one stream, a stream-ordered workspace allocator and a
// launch counter (bench.py reports it as gpu_launches).
struct Ctx {
    int device = 0;
    int pointer_mode = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    long long launches = 0;
    std::string last_error;
    double ccsd_loop_seconds = 0.0;
    const double *ccsd_t1_start = nullptr;// one-shot: device T1 the next CCSD loop starts from instead of zero (qcp2_ccsd_set_start_T1)
    int ccsd_diis = 0;      // opt-in DIIS subspace size of the CCSD loop (qcp2_set_ccsd_diis; 0 = the reference's plain Jacobi loop)
    int last_ccsd_mode = -1;// formulation the last CCSD solve used: 0 dense term list, 1 pair-packed, 2 pair-packed + spin classes
    PermCache *perm_cache = nullptr;
    // row-sharded GEMMs (dgemm.cu): while set, every product above shard_min_flops whose result is a dense row-major
    // matrix is split by rows over the communicator's ranks and re-assembled with one broadcast per rank
    struct Comm *shard = nullptr;
    double shard_min_flops = 8.0e9;// r2: the spin-blocked / packed products of real molecules are below it (sharding them cost 12 % at CH4/cc-pVTZ)
    long long sharded_gemms = 0;
    // split-K tile counters of the warp-specialised GEMM (dgemm_ws.cuh): zero between launches (the last CTA of a
    // tile resets its counter), allocated with the context
    unsigned int *ws_counters = nullptr;
    int ws_counters_cap = 0;
    long long ws_gemms = 0, legacy_gemms = 0;
    // optional record of the products gemm() launches (qcp2_gemm_log): bench.py derives the executed flops and the
    // operand bytes of one CCSD iteration from it
    struct GemmRec {
        int M, N, K, batch, beta, tile_m, tile_n, split;
    };
    std::vector<GemmRec> gemm_log;
    bool log_gemms = false;
    // device offset tables of the GEMM's scatter epilogue, keyed by the contraction's signature (contract.cu); they
    // live as long as the context, so a table made in an eager iteration is there when the same call is captured
    struct OffsetTable {
        std::string sig;
        long long *dev;
    };
    std::vector<OffsetTable> offset_tables;

    // workspace pool owned by this context (created in qcp2_create, trimmed and destroyed in qcp2_destroy): freed
    // blocks stay cached here instead of in the device's default pool, so the library does not change the behaviour
    // of other cudaMallocAsync users of the process
    cudaMemPool_t pool = nullptr;
    unsigned long long pool_reset_value = 0;// written to cudaMemPoolAttrUsedMemHigh to restart the high-water mark

    double *alloc(size_t n_doubles) { return static_cast<double *>(alloc_bytes((n_doubles ? n_doubles : 1) * sizeof(double))); }
    void *alloc_bytes(size_t n) {
        void *p = nullptr;
        if (pool) QCP2_CUDA(cudaMallocFromPoolAsync(&p, n ? n : 1, pool, stream));
        else QCP2_CUDA(cudaMallocAsync(&p, n ? n : 1, stream));
        return p;
    }
    void free(void *p) {
        if (p) cudaFreeAsync(p, stream);
    }
    void sync() { QCP2_CUDA(cudaStreamSynchronize(stream)); }
    void launched(const char *what) {
        ++launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) throw Error(std::string("kernel launch failed (") + what + "): " + cudaGetErrorString(e));
    }
};

// QCP2_TRACE=1: wall-clock phase marks (each mark synchronises the stream) on stderr
struct Trace {
    Ctx &c;
    const char *what;
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace(Ctx &ctx, const char *w) : c(ctx), what(w), on(getenv("QCP2_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *phase) {
        if (!on) return;
        c.sync();
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[qcp2 trace] %s: %-44s %9.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// RAII stream-ordered device buffer
struct DBuf {
    Ctx *ctx = nullptr;
    double *p = nullptr;
    size_t n = 0;
    DBuf() {}
    DBuf(Ctx &c, size_t n_) : ctx(&c), p(c.alloc(n_)), n(n_) {}
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; }
    DBuf &operator=(DBuf &&o) noexcept {
        if (this != &o) {
            reset();
            ctx = o.ctx, p = o.p, n = o.n;
            o.p = nullptr;
        }
        return *this;
    }
    void reset() {
        if (p && ctx) ctx->free(p);
        p = nullptr;
    }
    ~DBuf() { reset(); }
    operator double *() const { return p; }
};

// Dense row-major tensor view on device memory (rank <= 8).
struct TView {
    double *p = nullptr;
    int rank = 0;
    long long ext[8] = {1, 1, 1, 1, 1, 1, 1, 1};
    TView() {}
    TView(double *ptr, std::initializer_list<long long> e) : p(ptr), rank((int) e.size()) {
        int k = 0;
        for (long long x : e) ext[k++] = x;
    }
    TView(const double *ptr, std::initializer_list<long long> e) : TView(const_cast<double *>(ptr), e) {}
    long long size() const {
        long long n = 1;
        for (int k = 0; k < rank; ++k) n *= ext[k];
        return n;
    }
};

inline unsigned grid_for(long long n, int block, int sm_count, int per_sm = 16) {
    long long g = (n + block - 1) / block;
    long long cap = (long long) sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned) g;
}

}// namespace qcp2
