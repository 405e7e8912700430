// comm.cu -- run-time binding of NCCL and the collectives the hot path needs (SURVEY.md
// section 8e): broadcast of the replicated (T) inputs, all-reduce of the per-triple energy
// vector, all-gather for column-sharded GEMM results.
#include "comm.cuh"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace qcp2 {

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: stable enum values and signatures)
typedef void *nccl_comm_t;
struct nccl_unique_id {
    char internal[128];
};
enum { kNcclSuccess = 0, kNcclUint8 = 1, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        // the SONAME lookup returns a copy the process already loaded (e.g. the one bundled with
        // a Python host), otherwise the system library
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return;
        auto sym = [&](const char *n) { return dlsym(a.handle, n); };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    });
    if (!a.handle || !a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Broadcast || !a.AllReduce || !a.AllGather || !a.GroupStart ||
        !a.GroupEnd)
        throw Error("NCCL is not available (dlopen of libnccl.so.2 failed or symbols are missing)");
    return a;
}

void check(int rc, const char *what) {
    if (rc == kNcclSuccess) return;
    NcclApi &a = api();
    throw Error(std::string(what) + " failed: " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error " + std::to_string(rc)));
}

}// namespace

void comm_unique_id(void *id128) {
    nccl_unique_id id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, id.internal, sizeof(id.internal));
}

Comm *comm_init(Ctx &ctx, const void *id128, int nranks, int rank) {
    QCP2_REQUIRE(id128 && nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad arguments");
    QCP2_CUDA(cudaSetDevice(ctx.device));
    nccl_unique_id id;
    std::memcpy(id.internal, id128, sizeof(id.internal));
    Comm *c = new Comm();
    c->rank = rank, c->nranks = nranks;
    const int rc = api().CommInitRank(&c->nccl, nranks, id, rank);
    if (rc != kNcclSuccess) {
        delete c;
        check(rc, "ncclCommInitRank");
    }
    return c;
}

void comm_destroy(Comm *comm) {
    if (!comm) return;
    if (comm->nccl) api().CommDestroy(comm->nccl);
    delete comm;
}

void comm_broadcast(Ctx &ctx, Comm &comm, double *buf, long long n, int root, cudaStream_t stream) {
    if (n <= 0 || comm.nranks == 1) return;
    check(api().Broadcast(buf, buf, (size_t) n, kNcclFloat64, root, comm.nccl, stream ? stream : ctx.stream), "ncclBroadcast");
}

void comm_broadcast_bytes(Ctx &ctx, Comm &comm, void *buf, long long nbytes, int root, cudaStream_t stream) {
    if (nbytes <= 0 || comm.nranks == 1) return;
    check(api().Broadcast(buf, buf, (size_t) nbytes, kNcclUint8, root, comm.nccl, stream ? stream : ctx.stream), "ncclBroadcast");
}

void comm_allreduce_sum(Ctx &ctx, Comm &comm, double *buf, long long n) {
    if (n <= 0 || comm.nranks == 1) return;
    check(api().AllReduce(buf, buf, (size_t) n, kNcclFloat64, kNcclSum, comm.nccl, ctx.stream), "ncclAllReduce");
}

// Collective error agreement: every rank contributes 0 (fine) or 1 (local_err is not empty); the sum is read
// back and, if any rank failed, EVERY rank throws -- so a failure on one rank before a payload collective
// can no longer leave the others blocked in it (qcp2_ccsd_t_mgpu / qcp2_post_hf_mgpu hang).
void comm_agree(Ctx &ctx, Comm &comm, const std::string &local_err, const char *where, double *vote) {
    if (comm.nranks == 1) {
        if (!local_err.empty()) throw Error(local_err);
        return;
    }
    double *flag = ctx.alloc(2);
    const double mine[2] = {local_err.empty() ? 0.0 : 1.0, vote ? *vote : 0.0};
    double sums[2] = {0.0, 0.0};
    try {
        QCP2_CUDA(cudaMemcpyAsync(flag, mine, 16, cudaMemcpyHostToDevice, ctx.stream));
        comm_allreduce_sum(ctx, comm, flag, 2);
        QCP2_CUDA(cudaMemcpyAsync(sums, flag, 16, cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
    } catch (...) {
        ctx.free(flag);
        throw;
    }
    ctx.free(flag);
    const double total = sums[0];
    if (vote) *vote = sums[1];
    if (total != 0.0) {
        if (!local_err.empty()) throw Error(local_err);
        throw Error(std::string(where) + ": " + std::to_string((int) total) + " other rank(s) failed before the exchange; nothing was computed");
    }
}

void comm_group_start() { check(api().GroupStart(), "ncclGroupStart"); }
void comm_group_end() { check(api().GroupEnd(), "ncclGroupEnd"); }

void comm_allgather(Ctx &ctx, Comm &comm, const double *send, double *recv, long long n_per_rank, cudaStream_t stream) {
    if (n_per_rank <= 0) return;
    cudaStream_t st = stream ? stream : ctx.stream;
    if (comm.nranks == 1) {
        if (send != recv) QCP2_CUDA(cudaMemcpyAsync(recv, send, (size_t) n_per_rank * 8, cudaMemcpyDeviceToDevice, st));
        return;
    }
    check(api().AllGather(send, recv, (size_t) n_per_rank, kNcclFloat64, comm.nccl, st), "ncclAllGather");
}

}// namespace qcp2
