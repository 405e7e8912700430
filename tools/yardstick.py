"""FP64 roofline yardstick (MEASURED_PEAKS.json has bf16 + HBM only): cuBLAS DGEMM through
torch.matmul(float64) as burst (best of N) and sustained (back to back for ~4 s), next to this
repo's DMMA GEMM on the same shapes.  cuBLAS is used ONLY as the denominator, never on the
product path.  Writes gpurun_out/fp64_peak.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qcp2_b200  # noqa: E402


def time_loop(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def cublas_fp64_peak(n=8192, burst_reps=10, sustain_seconds=4.0):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    fl = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    best = min(time_loop(lambda: torch.matmul(a, b, out=c), 1) for _ in range(burst_reps))
    t1 = time_loop(lambda: torch.matmul(a, b, out=c), 5)
    iters = max(5, int(sustain_seconds / t1))
    sus = time_loop(lambda: torch.matmul(a, b, out=c), iters)
    return dict(n=n, burst_tflops=fl / best / 1e12, sustained_tflops=fl / sus / 1e12, sustained_iters=iters)


def main():
    out = {"device": torch.cuda.get_device_name(0)}
    out["cublas_dgemm"] = cublas_fp64_peak()
    print("cuBLAS", out["cublas_dgemm"], flush=True)
    ctx = qcp2_b200.Context(0)
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    rows = []
    shapes = [(8192, 8192, 8192, 0, 0), (8192, 8192, 8192, 0, 1), (8192, 8192, 8192, 1, 0),
              (400, 79800 * 4, 1344, 0, 0), (100, 26244, 26244, 0, 1), (26244, 26244, 100, 1, 0), (4096, 4096, 4096, 0, 0)]
    for (M, N, K, tA, tB) in shapes:
        A = torch.randn((K, M) if tA else (M, K), dtype=torch.float64, device="cuda")
        B = torch.randn((N, K) if tB else (K, N), dtype=torch.float64, device="cuda")
        Cd = torch.empty((M, N), dtype=torch.float64, device="cuda")
        fn = lambda: ctx.dgemm_device(tA, tB, M, N, K, 1.0, A, A.shape[1], B, B.shape[1], 0.0, Cd, N)
        for _ in range(2):
            fn()
        t = min(time_loop(fn, 1) for _ in range(5))
        ref = time_loop(lambda: torch.matmul(A.T if tA else A, B.T if tB else B, out=Cd), 3)
        row = dict(M=M, N=N, K=K, tA=tA, tB=tB, qcp2_tflops=2.0 * M * N * K / t / 1e12,
                   cublas_tflops=2.0 * M * N * K / ref / 1e12)
        rows.append(row)
        print(row, flush=True)
        del A, B, Cd
    out["qcp2_dgemm"] = rows
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fp64_peak.json"), "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
