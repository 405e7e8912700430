"""One transform_ao_mo at n = 86 (CH4/cc-pVTZ) on device tensors, for ncu: ncu --set full -k regex:ws_dgemm ... python tools/ao_mo_once.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import qcp2_b200

ctx = qcp2_b200.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
ctx.set_pointer_mode(qcp2_b200.DEVICE)
n = 86
f = dict(dtype=torch.float64, device="cuda")
ao, C1, out = torch.randn((n,) * 4, **f), torch.randn((n, n), **f), torch.empty((n,) * 4, **f)
p = qcp2_b200._p
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    torch.cuda.synchronize()
    e0.record()
    ctx.check(ctx.lib.qcp2_transform_ao_mo(ctx.h, p(ao), p(C1), p(C1), n, p(out)))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("transform_ao_mo n=%d: %.3f ms, %.2f TFLOP/s (8 n^5 flop)" % (n, ms, 8.0 * n ** 5 / ms / 1e9), flush=True)
ctx.close()
