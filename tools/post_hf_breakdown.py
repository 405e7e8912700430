"""Phase breakdown (QCP2_TRACE=1) of the fused MP2 -> CCSD -> (T) call on CH4/cc-pVTZ, cold and warm."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["QCP2_TRACE"] = "1"
import qcp2_b200  # noqa: E402
from helpers import molecule_case  # noqa: E402

t0 = time.perf_counter()
case = molecule_case("methane.xyz", "cc-pvtz", conv=1e-10, maxiter=400)
print("host SCF + AO integrals: %.2f s" % (time.perf_counter() - t0), flush=True)
ctx = qcp2_b200.Context(0)
for rep in range(2):
    t0 = time.perf_counter()
    r = ctx.post_hf(case["ao"], case["C"], None, case["eps"], None, case["no"], case["nv"], stages=3, maxiter=200, cc_conv=1e-12)
    print("   CCSD loop only: %.1f ms" % (1e3 * float(ctx.lib.qcp2_ccsd_last_loop_seconds(ctx.h))), flush=True)
    print("rep %d: %.3f s  E_CCSD %.12f  E_(T) %.12f  iters %d" % (rep, time.perf_counter() - t0, r["e_ccsd"], r["e_t"], r["iters"]), flush=True)
ctx.close()
