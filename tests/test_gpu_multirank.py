"""Multi-rank parity where the driver can see it (-m gpu): with >= 2 GPUs visible, two ranks (one process per GPU,
launched with torch.distributed.run) call the collective entry points qcp2_ccsd_t_mgpu and qcp2_post_hf_mgpu and the
energies must equal the one-rank results BIT FOR BIT (fixed triple -> slot map, one all-reduce in which every slot has a
single writer, fixed-order host sum; replicated CCSD with all-gathered sharded products).  On a one-GPU box the tests
skip with a visible reason -- two ranks on one GPU would be kernels waiting on one another, which the pool forbids."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["QCP2_ROOT"]); sys.path.insert(0, os.path.join(os.environ["QCP2_ROOT"], "tests"))
import qcp2_b200
from oracle import np_oracle as npo          # checker only
from helpers import rhf_case
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ctx = qcp2_b200.Context(local)
ctx.comm_init_from_torch(dist)
out = {}
names = ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")
for (o, v) in ((7, 30), (6, 33)):              # v(v-1)/2 even and odd
    s = npo.synthetic_t_inputs(o, v)
    e = ctx.CCSD_T_mgpu(*[s[k] if rank == 0 else None for k in names], o, v, qcp2_b200.T_AUTO, root=0)
    single = ctx.CCSD_T(*(s[k] for k in names), qcp2_b200.T_SYMMETRIC)
    box = [None] * world
    dist.all_gather_object(box, (e, single))
    out["t_o%dv%d" % (o, v)] = {"mgpu": e, "single": single, "all_ranks": box}
# replicated host inputs: EVERY rank passes the (identical) tensors -> the <vo||vv> slabs are dealt over the ranks' PCIe links
s = npo.synthetic_t_inputs(7, 30)
e_rep = ctx.CCSD_T_mgpu(*[s[k] for k in names], 7, 30, qcp2_b200.T_AUTO, root=0)
box = [None] * world
dist.all_gather_object(box, e_rep)
out["t_replicated_host"] = {"mgpu": e_rep, "single": out["t_o7v30"]["single"], "all_ranks": box}
case = rhf_case(8, 3, 5)                        # seeded RHF-type pseudo-molecule, n = 8 spatial orbitals
args = (case["ao"], case["C"], None, case["eps"], None) if rank == 0 else (None,) * 5
m = ctx.post_hf_mgpu(*args, case["ao"].shape[0], case["no"], case["nv"], maxiter=60, cc_conv=1e-12)
one = ctx.post_hf(case["ao"], case["C"], None, case["eps"], None, case["no"], case["nv"], maxiter=60, cc_conv=1e-12)
box = [None] * world
dist.all_gather_object(box, (m["e_mp2"], m["e_ccsd"], m["e_t"], m["iters"]))
out["post_hf"] = {"mgpu": box, "single": (one["e_mp2"], one["e_ccsd"], one["e_t"], one["iters"])}
ctx.close()
if rank == 0:
    print("RESULT " + json.dumps(out))
dist.destroy_process_group()
'''


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def two_rank_result(tmp_path_factory):
    n = _gpu_count()
    if n < 2:
        pytest.skip(f"multi-rank parity needs >= 2 GPUs on the box, {n} visible (run: gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu)")
    script = tmp_path_factory.mktemp("mr") / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, QCP2_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_two_rank_triples_energy_is_bit_identical(two_rank_result):
    for key in ("t_o7v30", "t_o6v33"):
        rec = two_rank_result[key]
        assert rec["mgpu"] == rec["single"], rec                       # bit for bit
        assert all(tuple(b) == (rec["mgpu"], rec["single"]) for b in rec["all_ranks"]), rec


def test_two_rank_dealt_upload_is_bit_identical(two_rank_result):
    rec = two_rank_result["t_replicated_host"]
    assert rec["mgpu"] == rec["single"] and all(b == rec["mgpu"] for b in rec["all_ranks"]), rec


def test_two_rank_post_hf_matches_one_rank(two_rank_result):
    rec = two_rank_result["post_hf"]
    e2, ecc, et, iters = rec["single"]
    for b in rec["mgpu"]:
        assert b[3] == iters
        assert b[0] == e2                                              # replicated MP2: identical arithmetic
        assert abs(b[1] - ecc) < 1e-12 and abs(b[2] - et) < 1e-13, (b, rec["single"])
    assert all(b == rec["mgpu"][0] for b in rec["mgpu"])               # every rank returns the same numbers
