"""Multi-rank check of the fused, collective pipeline qcp2_post_hf_mgpu (run under torchrun on N GPUs of one box):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_ccsd_check.py

Rank 0 runs the host SCF of CH4/cc-pVTZ (o=10, v=162) and of water/cc-pVDZ and supplies the boundary inputs; every rank joins
qcp2_post_hf_mgpu (row-sharded CCSD products + round-robin (T)).  Checks: identical energies on every rank, equal to
the single-GPU qcp2_post_hf of rank 0 to 1e-10 with the same iteration count; reports the CCSD loop time per iteration
sharded vs replicated (QCP2_CCSD_NO_SHARD=1) vs single GPU.  Writes gpurun_out/mgpu_ccsd_check_N.json on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qcp2_b200  # noqa: E402
from helpers import molecule_case  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = qcp2_b200.Context(local)
    ctx.comm_init_from_torch(dist)
    out = {"world": world}
    for xyz, basis, kw in (("water.xyz", "cc-pvdz", {}), ("methane.xyz", "cc-pvtz", dict(conv=1e-10, maxiter=400))):
        case = molecule_case(xyz, basis, **kw) if rank == 0 else None
        dims = [(case["n"], case["no"], case["nv"]) if rank == 0 else None]
        dist.broadcast_object_list(dims, src=0)
        n, no, nv = dims[0]
        args = (case["ao"], case["C"], None, case["eps"], None) if rank == 0 else (None,) * 5
        res = {}
        for label, env in (("sharded", None), ("replicated", "1")):
            if env:
                os.environ["QCP2_CCSD_NO_SHARD"] = env
            else:
                os.environ.pop("QCP2_CCSD_NO_SHARD", None)
            for rep in range(2):  # second call is warm
                g0 = ctx.sharded_gemms
                r = ctx.post_hf_mgpu(*args, n, no, nv, stages=3, maxiter=200, cc_conv=1e-12, root=0)
            loop = float(ctx.lib.qcp2_ccsd_last_loop_seconds(ctx.h))
            res[label] = dict(e_mp2=r["e_mp2"], e_ccsd=r["e_ccsd"], e_t=r["e_t"], iters=r["iters"], ms_per_iter=1e3 * loop / max(r["iters"], 1),
                              sharded_products_per_call=ctx.sharded_gemms - g0)
            box = [None] * world
            dist.all_gather_object(box, (r["e_mp2"], r["e_ccsd"], r["e_t"], r["iters"]))
            assert all(b == box[0] for b in box), box  # every rank: the same bits
        os.environ.pop("QCP2_CCSD_NO_SHARD", None)
        if rank == 0:
            for rep in range(2):
                s = ctx.post_hf(case["ao"], case["C"], None, case["eps"], None, no, nv, stages=3, maxiter=200, cc_conv=1e-12)
            loop = float(ctx.lib.qcp2_ccsd_last_loop_seconds(ctx.h))
            res["single_gpu"] = dict(e_mp2=s["e_mp2"], e_ccsd=s["e_ccsd"], e_t=s["e_t"], iters=s["iters"], ms_per_iter=1e3 * loop / max(s["iters"], 1))
            for label in ("sharded", "replicated"):
                for k in ("e_mp2", "e_ccsd", "e_t"):
                    assert abs(res[label][k] - s[k]) < 1e-10, (label, k, res[label][k], s[k])
                assert res[label]["iters"] == s["iters"]
            out[f"{xyz}/{basis}"] = res
            print(xyz, basis, json.dumps(res), flush=True)
        dist.barrier()
    ctx.close()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"mgpu_ccsd_check_{world}.json"), "w"), indent=1)
        print("mgpu_ccsd_check OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
