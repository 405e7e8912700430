"""CCSD seconds/iteration at the CH4/cc-pVTZ shape (o=10, v=162), alone (the secondary metric of bench.py),
for profiling:  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file ... python tools/ccsd_iter.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import qcp2_b200  # noqa: E402

ctx = qcp2_b200.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
iters = int(os.environ.get("QCP2_CCSD_ITERS", "3"))
print(json.dumps(bench.ccsd_seconds_per_iter(torch, qcp2_b200, ctx, iters=iters)))
ctx.close()
