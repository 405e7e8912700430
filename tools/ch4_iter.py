"""CH4/cc-pVTZ (BASELINE.json config 3) through the fused qcp2_post_hf, CCSD only, a few iterations: for ncu launch lists of
the REAL-molecule iteration (spin-blocked pairs).  QCP2_CCSD_NO_GRAPH=1 QCP2_CH4_ITERS=3 ncu ... python tools/ch4_iter.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qcp2_b200  # noqa: E402
from helpers import molecule_case  # noqa: E402

case = molecule_case("methane.xyz", "cc-pvtz", conv=1e-10, maxiter=400)
ctx = qcp2_b200.Context(0)
iters = int(os.environ.get("QCP2_CH4_ITERS", "3"))
for rep in range(2):
    r = ctx.post_hf(case["ao"], case["C"], None, case["eps"], None, case["no"], case["nv"], stages=2, maxiter=iters, cc_conv=0.0)
    print("rep %d mode %d: CCSD loop %.3f ms / iteration, E after %d iterations %.12f" % (
        rep, ctx.ccsd_last_mode(), 1e3 * float(ctx.lib.qcp2_ccsd_last_loop_seconds(ctx.h)) / iters, iters, r["e_ccsd"]), flush=True)
ctx.close()
