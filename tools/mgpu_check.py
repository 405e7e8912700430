"""Multi-rank check of the library's own NCCL path (run under torchrun on N GPUs of one box):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py

Every rank creates a context on its GPU, the ranks bootstrap qcp2_comm_init through torch.distributed (gloo; it only
carries the 128-byte id) and call qcp2_ccsd_t_mgpu on a reduced synthetic shape with the tensors supplied by rank 0
only.  Checks: the energy is identical on every rank, bit-identical to the single-GPU qcp2_ccsd_t of rank 0, and equal
to the CPU oracle to 1e-10; broadcast / all-reduce / all-gather building blocks behave.  Writes
gpurun_out/mgpu_check_N.json on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qcp2_b200  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402  (checker only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = qcp2_b200.Context(local)
    ctx.comm_init_from_torch(dist)
    assert ctx.comm_rank() == (rank, world)
    names = ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")
    out = {"world": world}
    for (o, v) in ((7, 30), (12, 96)):
        s = npo.synthetic_t_inputs(o, v)
        src = [s[k] if rank == 0 else None for k in names]
        e, per = ctx.CCSD_T_mgpu(*src, o, v, qcp2_b200.T_AUTO, root=0, want_triples=True)
        single = ctx.CCSD_T(*(s[k] for k in names), qcp2_b200.T_SYMMETRIC)  # every rank, on its own copy
        cpu = npo.ccsd_t_blocked(*(s[k] for k in names), True) if rank == 0 and o == 7 else None
        box = [None] * world
        dist.all_gather_object(box, (e, single))
        assert all(b[0] == box[0][0] for b in box), box           # same answer on every rank
        assert e == single or abs(e - single) <= 4e-16 * abs(single), (e, single)
        out[f"o{o}v{v}"] = dict(e_mgpu=e, e_single=single, bit_identical=(e == single), e_cpu=cpu, ntriples=int(np.count_nonzero(per)))
        if cpu is not None:
            assert abs(e - cpu) < 1e-10 * max(1.0, abs(cpu))
    # literal mode (unsymmetric inputs) across ranks
    rng = np.random.default_rng(3)
    o, v = 4, 5
    lit = [rng.standard_normal(sh) * 0.1 for sh in ((o, v), (o, o, v, v), (o, o, v, v), (v, o, v, v), (o, v, o, o))]
    lit += [-1.0 - rng.random(o), 1.0 + rng.random(v)]
    e_lit = ctx.CCSD_T_mgpu(*[a if rank == 0 else None for a in lit], o, v, qcp2_b200.T_AUTO, root=0)
    assert e_lit == ctx.CCSD_T(*lit) or abs(e_lit - ctx.CCSD_T(*lit)) < 1e-15
    out["literal"] = e_lit
    # building blocks
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    x = torch.full((1000,), float(rank + 1), dtype=torch.float64, device="cuda")
    ctx.comm_allreduce_sum(x)
    g = torch.empty(world * 3, dtype=torch.float64, device="cuda")
    ctx.comm_allgather(torch.full((3,), float(rank), dtype=torch.float64, device="cuda"), g)
    b = torch.full((7,), float(rank), dtype=torch.float64, device="cuda")
    ctx.comm_broadcast(b, root=world - 1)
    ctx.synchronize()
    assert float(x[0]) == world * (world + 1) / 2 and float(b[0]) == world - 1
    assert g.cpu().tolist() == [float(r) for r in range(world) for _ in range(3)]
    ctx.close()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"mgpu_check_{world}.json"), "w"), indent=1)
        print("mgpu_check OK", json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
