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
    print("rep %d: %.3f s  E_CCSD %.12f  E_(T) %.12f  iters %d  peak workspace %.3f GB" % (rep, time.perf_counter() - t0, r["e_ccsd"], r["e_t"], r["iters"], ctx.pool_high_water() / 1e9), flush=True)
if os.environ.get("QCP2_LARGE_N"):
    # synthetic n = 150 (o = 10, v = 290): the (2n)^4 tensor would be 65 GB and <vv||vv> 56 GB; neither is built
    import numpy as np
    import torch
    n, ndocc = int(os.environ["QCP2_LARGE_N"]), 5
    g = torch.Generator(device="cuda").manual_seed(11)
    ao = torch.randn((n,) * 4, generator=g, dtype=torch.float64, device="cuda")
    ao = ao + ao.permute(1, 0, 2, 3)
    ao = ao + ao.permute(0, 1, 3, 2)
    ao = ((ao + ao.permute(2, 3, 0, 1)) * (0.02 / 8.0)).contiguous()
    from helpers import fake_orbitals
    Cm, eps = fake_orbitals(n, ndocc, 3)
    Cd, ed = torch.tensor(Cm, device="cuda"), torch.tensor(eps, device="cuda")
    torch.cuda.synchronize()
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    import ctypes as C
    e2, ecc, et, iters = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    p = qcp2_b200._p
    t0 = time.perf_counter()
    ctx.check(ctx.lib.qcp2_post_hf(ctx.h, p(ao), p(Cd), None, p(ed), None, n, 2 * ndocc, 2 * (n - ndocc), 0, 3, 12, C.c_double(1e-9),
                                   C.byref(e2), C.byref(ecc), C.byref(et), C.byref(iters), None, None, None))
    print("synthetic n=%d (o=%d, v=%d): %.2f s, E_MP2 %.10f E_CCSD %.10f E_(T) %.3e, %d iterations, CCSD loop %.1f ms/iter, peak workspace %.2f GB"
          % (n, 2 * ndocc, 2 * (n - ndocc), time.perf_counter() - t0, e2.value, ecc.value, et.value, iters.value,
             1e3 * float(ctx.lib.qcp2_ccsd_last_loop_seconds(ctx.h)) / max(iters.value, 1), ctx.pool_high_water() / 1e9), flush=True)
ctx.close()
