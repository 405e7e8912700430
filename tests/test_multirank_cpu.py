"""World-size-2 gloo test of the (T) sharding logic (qc-p2_b200/shard.py): static round-robin
of triples, zero-padded per-triple vector, one all-reduce, fixed-order sum.  The per-triple
numbers come from the numpy oracle here; on the GPU box the same driver code is fed by the
CUDA path (tests/test_gpu_parity.py, bench.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, o, v, out):
    sys.path.insert(0, ROOT)
    import qcp2_b200
    from oracle import np_oracle as npo
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = npo.synthetic_t_inputs(o, v) if rank == 0 else {k: np.zeros_like(x) for k, x in npo.synthetic_t_inputs(o, v).items()}
    tensors = {k: torch.from_numpy(np.ascontiguousarray(x)) for k, x in s.items()}
    qcp2_b200.shard.broadcast_inputs(list(tensors.values()), dist, src=0)
    s = {k: t.numpy() for k, t in tensors.items()}
    trip = npo.sorted_triples(o)
    e = torch.zeros(len(trip), dtype=torch.float64)
    first, stride = qcp2_b200.shard.rank_range(rank, world)
    n_mine = 0
    for t in range(first, len(trip), stride):
        i, j, k = map(int, trip[t])
        e[t] = npo.ccsd_t_triple_packed_np(s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"], i, j, k)
        n_mine += 1
    assert n_mine == qcp2_b200.shard.rank_count(len(trip), rank, world)
    total = qcp2_b200.shard.reduce_triple_energies(e, dist)
    out[rank] = total
    dist.destroy_process_group()


def test_two_rank_sharded_t_matches_single_rank_bitwise():
    from oracle import np_oracle as npo
    o, v = 5, 6
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, o, v, out), nprocs=2, join=True)
    s = npo.synthetic_t_inputs(o, v)
    trip = npo.sorted_triples(o)
    single = 0.0
    for i, j, k in trip:
        single += npo.ccsd_t_triple_packed_np(s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"],
                                              int(i), int(j), int(k))
    assert out[0] == out[1] == single  # bit-identical, not merely close
    full = npo.ccsd_t_blocked(s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"], False)
    assert abs(single - full) < 1e-13 * max(1.0, abs(full))


def test_rank_partition_covers_every_triple_once():
    import qcp2_b200
    for n in (0, 1, 7, 120, 9880):
        for world in (1, 2, 4, 8):
            seen = np.zeros(n, dtype=int)
            for r in range(world):
                first, stride = qcp2_b200.shard.rank_range(r, world)
                seen[first::stride] += 1
                assert qcp2_b200.shard.rank_count(n, r, world) == len(range(first, n, stride))
            assert np.all(seen == 1)


def _bootstrap_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import qcp2_b200
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the host half of Context.comm_init_from_torch: rank 0 makes the NCCL id, the process group hands it round
    box = [qcp2_b200.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    out[rank] = bytes(box[0])
    dist.destroy_process_group()


def test_two_rank_bootstrap_of_the_library_communicator_id():
    """qcp2_comm_get_unique_id needs no GPU: the 128-byte NCCL bootstrap id made by rank 0 reaches every rank unchanged
    (qcp2_comm_init itself is collective over GPUs and is exercised by tools/mgpu_check.py on the GPU box)."""
    import ctypes

    import qcp2_b200
    try:
        qcp2_b200.comm_unique_id()
    except qcp2_b200.QCP2Error:
        import pytest
        pytest.skip("libnccl.so.2 is not loadable here")
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_bootstrap_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == out[1] and len(out[0]) == qcp2_b200.COMM_ID_BYTES and any(out[0])
    # a too-small buffer is refused instead of overrun
    small = ctypes.create_string_buffer(16)
    assert qcp2_b200.load_library().qcp2_comm_get_unique_id(small, 16) != 0
