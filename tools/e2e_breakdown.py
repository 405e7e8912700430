"""Where the end-to-end (host-buffer) (T) call spends its time: QCP2_TRACE=1 phase marks of qcp2_ccsd_t on the
o=40/v=400 synthetic job restricted to a few triples (the run phase scales with the triple count, the rest does not)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["QCP2_TRACE"] = "1"
import qcp2_b200  # noqa: E402

ctx = qcp2_b200.Context(0)
o, v = 40, 400
d = qcp2_b200.synth_t_inputs_device(ctx, o, v)
names = ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")
host = {}
for k in names:
    h = torch.empty(d[k].shape, dtype=torch.float64, pin_memory=True)
    h.copy_(d[k])
    host[k] = h
del d
torch.cuda.empty_cache()
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    e = ctx.CCSD_T_range(*(host[k].numpy() for k in names), qcp2_b200.T_SYMMETRIC, 0, 1, 40)
    print("rep", rep, "40 triples, host buffers:", time.perf_counter() - t0, "s  E =", e, flush=True)
# raw pinned H2D rate for reference
x = torch.empty(host["vovv"].shape, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
x.copy_(host["vovv"], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("pinned H2D of vovv: %.3f s, %.1f GB/s" % (dt, x.numel() * 8 / dt / 1e9))
ctx.close()
