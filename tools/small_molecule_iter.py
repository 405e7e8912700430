"""CCSD loop time per iteration on the small real-molecule configs (launch-latency bound): CUDA-graph replay vs eager."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qcp2_b200  # noqa: E402
from helpers import molecule_case  # noqa: E402

ctx = qcp2_b200.Context(0)
for xyz, basis in (("water.xyz", "sto-3g"), ("water.xyz", "cc-pvdz")):
    case = molecule_case(xyz, basis)
    for rep in range(2):
        r = ctx.post_hf(case["ao"], case["C"], None, case["eps"], None, case["no"], case["nv"], stages=3, maxiter=200, cc_conv=1e-12)
    loop = float(ctx.lib.qcp2_ccsd_last_loop_seconds(ctx.h))
    print(f"{xyz}/{basis}: graph={'QCP2_CCSD_NO_GRAPH' not in os.environ}  {r['iters']} iterations, loop {1e3 * loop:.2f} ms, "
          f"{1e3 * loop / max(r['iters'], 1):.3f} ms/iteration, E_CCSD {r['e_ccsd']:.12f}", flush=True)
ctx.close()
