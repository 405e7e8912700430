"""Times the warp-specialised TMA/DMMA GEMM (qcp2_dgemm_ex) against cuBLAS DGEMM (torch.matmul) on the shapes that
matter: 8192^3, the packed particle-particle ladder (45 x 13041 x 13041), the (ov)^3 ring product, tau.<am||ef>.
usage: python tools/ws_gemm_bench.py > gpurun_out/ws_gemm_bench.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import qcp2_b200


def cublas(a, b, reps):
    torch.matmul(a, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    ctx = qcp2_b200.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    f = dict(dtype=torch.float64, device="cuda")
    out = []
    cases = [("8192^3 NN", 8192, 8192, 8192, 0, 0, [(-1, 0), (0, 0), (4, 0)]),
             ("8192^3 NT", 8192, 8192, 8192, 0, 1, [(-1, 0), (0, 0)]),
             ("8192^3 TN", 8192, 8192, 8192, 1, 0, [(0, 0)]),
             ("packed ladder 45x13041x13041 NT", 45, 13041, 13041, 0, 1, [(-1, 0), (1, 1), (1, 2), (1, 3), (1, 4), (1, 6), (1, 8)]),
             ("ring (ov)^3 1620^3 NN", 1620, 1620, 1620, 0, 0, [(-1, 0), (0, 1), (4, 1), (0, 2), (4, 2)]),
             ("tau.<am||ef> packed 45x1620x13041 NT", 45, 1620, 13041, 0, 1, [(-1, 0), (1, 8), (1, 16), (1, 21)]),
             ("tau.<am||ef> dense 100x1620x26244 NT", 100, 1620, 26244, 0, 1, [(-1, 0), (4, 11)]),
             ("skinny 10x262440x162 NT", 10, 262440, 162, 0, 1, [(-1, 0), (2, 1)]),
             ("skinny 262440x10x162 NN", 262440, 10, 162, 0, 0, [(-1, 0), (18, 1)]),
             ("AO->MO 86x636056x86 TT", 86, 636056, 86, 1, 1, [(-1, 0), (4, 1), (0, 1), (7, 1)])]
    for name, M, N, K, tA, tB, variants in cases:
        sa = (K, M) if tA else (M, K)
        sb = (N, K) if tB else (K, N)
        lda, ldb, ldc = sa[1] + (sa[1] & 1), sb[1] + (sb[1] & 1), N + (N & 1)
        A = torch.randn((sa[0], lda), **f)
        B = torch.randn((sb[0], ldb), **f)
        Cm = torch.zeros((M, ldc), **f)
        a = A[:, :sa[1]].t() if tA else A[:, :sa[1]]
        b = B[:, :sb[1]].t() if tB else B[:, :sb[1]]
        flops = 2.0 * M * N * K
        reps = 3 if flops > 5e11 else 20
        t_cublas = cublas(a, b, reps)
        want = torch.matmul(a, b)
        rec = {"case": name, "M": M, "N": N, "K": K, "cublas_us": t_cublas * 1e6, "cublas_tflops": flops / t_cublas / 1e12, "ws": []}
        for tile, split in variants:
            try:
                ctx.dgemm_ex(tA, tB, M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, tile=tile, k_splits=split, reps=2)
                t = ctx.dgemm_ex(tA, tB, M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, tile=tile, k_splits=split, reps=reps)
                err = float((Cm[:, :N] - want).abs().max())
                rec["ws"].append({"tile": tile, "split": split, "us": t * 1e6, "tflops": flops / t / 1e12, "vs_cublas": t_cublas / t,
                                  "max_abs_err": err})
            except Exception as ex:
                rec["ws"].append({"tile": tile, "split": split, "error": str(ex)[:200]})
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del A, B, Cm, want
    ctx.close()


if __name__ == "__main__":
    main()
