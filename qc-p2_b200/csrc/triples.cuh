// triples.cuh -- perturbative triples (T), cc.cpp:454-560, evaluated one occupied triple at a
// time: one segmented-K DMMA GEMM X = A.B (M = v, K = 3(v+o)) per triple plus a fused
// permute-combine / denominator / energy sweep.  No o^3 v^3 tensor is ever materialised.
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "dgemm.cuh"

namespace qcp2 {

// ---- TMA-fed warp-specialised GEMM of the symmetric path (t_gemm_tma.cu) ----
constexpr int TG_BM = 80, TG_BN = 256, TG_STAGES = 4, TG_ALD = GEMM_BK + 4;
struct TmaGemmArgs {
    const double *At = nullptr;  // pre-tiled A: [triple][m-tile][k-block][TG_BM][TG_ALD]
    long long strideA = 0;       // doubles per triple
    double *C = nullptr;
    long long ldc = 0, strideC = 0;
    int M = 0, N = 0, nkb = 0;
    const SegDesc *seg = nullptr;
    int nseg = 0;
    int seg_kb_begin[GEMM_MAX_SEG + 1] = {0, 0, 0, 0, 0, 0, 0};
    int seg_rows[GEMM_MAX_SEG] = {0, 0, 0, 0, 0, 0};
    long long seg_ldb[GEMM_MAX_SEG] = {0, 0, 0, 0, 0, 0};
    int merged_sub_rows = 0;     // > 0: segment 3 is the merged hole block, rows [q s, (q+1) s) come from seg->B[3 + q]
};
void launch_t_gemm_tma(Ctx &ctx, const TmaGemmArgs &g, int batch);

struct Comm;
// where a streamed plan gets <vo||vv> from: the root rank's HOST tensor [v][o][v][v] (null on the
// other ranks), broadcast slab by slab when a communicator with more than one rank is given
struct TStream {
    const double *vovv_host = nullptr;
    Comm *comm = nullptr;
    int root = 0;
    // every rank holds the same host tensor (e.g. one shared-memory copy): slab i is uploaded and packed by rank i % W over its
    // own PCIe link and broadcast from there, so the 20 GB of H2D traffic are spread over W links instead of the root's one
    bool dealt = false;
};

struct TPlan {
    Ctx *ctx = nullptr;
    int o = 0, v = 0, mode = 0;     // mode: QCP2_T_LITERAL or QCP2_T_SYMMETRIC
    long long ncols = 0;            // v*v (literal) or v(v-1)/2 (symmetric, b<c packed)
    long long ld = 0;               // row stride of the re-laid tensors and of X: ncols rounded up to even
    long long ntriples = 0;         // o^3 or C(o,3)
    int kpad = 0;                   // padded K of the per-triple GEMM
    int kb_begin[GEMM_MAX_SEG + 1] = {0, 0, 0, 0, 0, 0, 0};
    // inputs (device).  Both modes own re-laid copies of <vo||vv> and T2 with contiguous, 16-byte aligned B rows
    // (symmetric: b<c packed; literal: every (b,c)); T1, <oo||vv> (literal) and <ov||oo> are read in place.
    const double *T1 = nullptr, *T2 = nullptr, *oovv = nullptr, *vovv = nullptr, *ovoo = nullptr;
    DBuf eo, ev, vovvP, T2P, oovvP;
    // double-buffered per-batch workspace
    int tb = 1;                     // triples per batch
    DBuf Abuf[2], Xbuf[2], partial[2], ebatch[2];
    SegDesc *seg_dev = nullptr;     // descriptors of the current run
    int *ijk_dev = nullptr;
    long long run_capacity = 0;
    cudaStream_t aux = nullptr;
    cudaEvent_t gemm_done[2] = {nullptr, nullptr}, energy_done[2] = {nullptr, nullptr};
    cudaEvent_t t_begin = nullptr, t_end = nullptr;
    std::vector<cudaEvent_t> gemm_ev;// begin/end pairs around every GEMM launch of the last run
    double last_gemm_seconds = 0.0, last_total_seconds = 0.0;
    long long last_gemm_launches = 0, last_all_launches = 0;
    int blocks_per_triple = 0;
    int sym_items = 0;              // work items (a, bt <= ct) of the symmetric energy sweep
    int *items_dev = nullptr;
    // streamed <vo||vv> (host-buffer entry points): the unpacked tensor never exists on the device; slab
    // vovv[:,i,:,:] is copied H2D into `stage`, packed into vovvP[i] (and, with a communicator, broadcast)
    // on the `copy` stream while the triples whose slabs have arrived are already being evaluated
    cudaStream_t copy = nullptr;
    std::vector<cudaEvent_t> slab_ready;  // one per occupied index
    DBuf stage;
    unsigned long long *vovv_defect = nullptr;// device [2]: max |x[e,b,c] + x[e,c,b]|, max |x| over the slabs this rank packed
    DBuf defect_all;                // dealt upload: the [2]-records of every rank (all-gathered)
    int defect_ranks = 0;
    // Streamed plans: the slab pipeline (upload + pack on the owning rank, ncclBroadcast, event) is enqueued by a helper
    // thread.  With a queue of NCCL calls behind the first uploads the host blocks inside ncclBroadcast until the uploads
    // drain (8 ranks: 0.37 s); on a helper thread that wait no longer keeps the main thread from launching the triples whose
    // slabs have arrived.  The main thread waits (host side) until the slab it needs has been ENQUEUED -- a
    // cudaStreamWaitEvent on a not yet recorded event would be a no-op -- and joins the thread before it touches the
    // communicator again.
    TStream src;
    bool src_multi = false, src_uploads = false;
    std::thread slab_thread;
    std::mutex slab_mutex;
    std::condition_variable slab_cv;
    long long slabs_enqueued = 0;   // guarded by slab_mutex
    std::string slab_error;         // guarded by slab_mutex
    long long slab_launches = 0;
    bool merged_o = false;          // the three hole segments share one padded K block (TMA kernel only)
    bool use_tma = false;           // the TMA/mbarrier kernel (t_gemm_tma.cu) runs the triples GEMM (both modes)
    long long a_per_triple = 0;     // doubles of A per triple (tiled or plain layout)
    ~TPlan();
};

// vovv is a device pointer, or null together with a TStream (SYMMETRIC mode only)
TPlan *t_plan_create(Ctx &ctx, const double *T1, const double *T2, const double *oovv, const double *vovv,
                     const double *ovoo, const double *eps_o, const double *eps_v, int o, int v, int mode,
                     const TStream *ts = nullptr);
// streamed plans: waits for the last slab and returns the antisymmetry defect of vovv in (b,c) seen by
// the rank that packed it (after a broadcast: by every rank)
void t_plan_vovv_defect(TPlan &plan, double *defect, double *maxabs);
void t_plan_wait_slab(TPlan &plan, long long i);   // host wait until slab i has been enqueued (throws the thread's error)
void t_plan_join_slabs(TPlan &plan);               // all slabs enqueued, helper thread gone
// evaluates triples first, first+stride, ... (< ntriples, at most max_count of them);
// e_triples_dev (device, ntriples doubles, may be null) receives each triple's weighted
// contribution at its global index; returns the partial sum in *energy_host.
void t_plan_run(TPlan &plan, long long first, long long stride, long long max_count, double *e_triples_dev,
                double *energy_host);
long long t_count(int o, int mode);
int t_select_mode(Ctx &ctx, const double *T2, const double *oovv, const double *vovv, const double *ovoo, int o, int v);

void synth_t_inputs(Ctx &ctx, int o, int v, unsigned long long seed, double *T1, double *T2, double *oovv, double *vovv,
                    double *ovoo, double *eps_o, double *eps_v);

}// namespace qcp2
