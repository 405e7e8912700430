// comm.cuh -- NCCL communicator of a context (one process / rank per GPU, NVLink 5 / NVSwitch).
// NCCL is bound at run time (dlopen of libnccl.so.2), so the library has no link-time
// dependency on it and shares the NCCL a host process may already have loaded.
#pragma once
#include "common.cuh"

namespace qcp2 {

struct Comm {
    void *nccl = nullptr;// ncclComm_t
    int rank = 0, nranks = 1;
};

void comm_unique_id(void *id128);
Comm *comm_init(Ctx &ctx, const void *id128, int nranks, int rank);
void comm_destroy(Comm *comm);
// stream-ordered collectives on device buffers of doubles / raw bytes
// (`stream`: enqueue there instead of on the context's stream -- the slab thread of a streamed (T) plan)
void comm_broadcast(Ctx &ctx, Comm &comm, double *buf, long long n, int root, cudaStream_t stream = nullptr);
void comm_broadcast_bytes(Ctx &ctx, Comm &comm, void *buf, long long nbytes, int root, cudaStream_t stream = nullptr);
void comm_allreduce_sum(Ctx &ctx, Comm &comm, double *buf, long long n);
// all ranks throw if any rank's local_err is non-empty (one 8-byte all-reduce); see comm.cu
// `vote` (optional): every rank's value is summed in the same exchange and the sum returned in place
void comm_agree(Ctx &ctx, Comm &comm, const std::string &local_err, const char *where, double *vote = nullptr);
// several collectives issued as one NCCL group (e.g. one broadcast per rank with unequal counts)
void comm_group_start();
void comm_group_end();
void comm_allgather(Ctx &ctx, Comm &comm, const double *send, double *recv, long long n_per_rank, cudaStream_t stream = nullptr);

}// namespace qcp2
