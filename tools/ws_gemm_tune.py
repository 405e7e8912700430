"""Tile / split-K sweep of the warp-specialised GEMM on the medium products of the spin-blocked CCSD iteration.
usage: python tools/ws_gemm_tune.py > gpurun_out/ws_gemm_tune.json"""
import json
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
f = dict(dtype=torch.float64, device="cuda")
cases = [("ring block 810^3 NN", 810, 810, 810, 0, 0, [(0, 1), (0, 2), (0, 3), (4, 1), (4, 2), (4, 3), (4, 4), (3, 1), (3, 2)]),
         ("ring block 405^3 NN", 405, 405, 405, 0, 0, [(0, 1), (0, 4), (4, 1), (4, 4), (4, 8), (3, 1), (3, 2), (3, 4)]),
         ("ladder mixed class 25x7396x6561 NT", 25, 7396, 6561, 0, 1, [(6, 1), (6, 3), (6, 5), (6, 10), (1, 5), (2, 5), (2, 10), (6, 15), (6, 20)]),
         ("ladder aa class 10x3656x3240 NT", 10, 3656, 3240, 0, 1, [(2, 1), (2, 5), (2, 10), (2, 20), (5, 10), (5, 15)]),
         ("U_ab 16200x162x162 NT", 16200, 162, 162, 0, 1, [(4, 1), (0, 1), (3, 1)]),
         ("Fae 162x162x16200 TN", 162, 162, 16200, 1, 0, [(3, 16), (3, 32), (3, 8), (4, 37), (4, 74), (0, 74), (0, 148)])]
for name, M, N, K, tA, tB, variants in cases:
    sa = (K, M) if tA else (M, K)
    sb = (N, K) if tB else (K, N)
    lda, ldb, ldc = sa[1] + (sa[1] & 1), sb[1] + (sb[1] & 1), N + (N & 1)
    A, B, Cm = torch.randn((sa[0], lda), **f), torch.randn((sb[0], ldb), **f), torch.zeros((M, ldc), **f)
    flops = 2.0 * M * N * K
    rec = {"case": name, "ws": []}
    ctx.dgemm_ex(tA, tB, M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, tile=-1, reps=3)
    t = ctx.dgemm_ex(tA, tB, M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, tile=-1, reps=20)
    rec["auto_us"] = t * 1e6
    for tile, split in variants:
        ctx.dgemm_ex(tA, tB, M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, tile=tile, k_splits=split, reps=3)
        t = ctx.dgemm_ex(tA, tB, M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, tile=tile, k_splits=split, reps=20)
        rec["ws"].append({"tile": tile, "split": split, "us": round(t * 1e6, 1), "tflops": round(flops / t / 1e12, 2)})
    print(json.dumps(rec), flush=True)
ctx.close()
