#!/usr/bin/env python
"""bench.py -- headline benchmark of the post-HF hot path on B200.

Metric (BASELINE.json): (T) FP64 TFLOP/s on the synthetic closed-shell shape o=40, v=400
(config 5; antisymmetrised counter-based random inputs), counted with the
ALGORITHMIC flop figure F_T = 2 * v * v(v-1)/2 * 3(v+o) per i<j<k triple, plus CCSD seconds per
iteration as a secondary object.

A "step" is one batch of `--triples-per-rank` occupied triples per GPU (weak scaling: per-GPU
work is fixed; 8 GPUs x 5 default steps = the whole 9 880-triple job).  Triples are dealt
round-robin over ranks; the only exchange is one all-reduce of the per-triple energy vector per
step.  `value` is timed with the inputs resident in HBM; `e2e` is the whole (T) job through the
reference-facing C-ABI call with HOST buffers (pinned), host->device copies, NCCL broadcast and
the final all-reduce inside the timed region.

  python bench.py [--gpus N --steps K --warmup W]           # our arm
  python bench.py --impl reference [...]                    # the CPU path on the host cores
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# The CPU legs (--impl reference, cpu_baseline) must use every host core.  torch.distributed.run exports
# OMP_NUM_THREADS=1 to its workers, and OpenBLAS / libgomp read the variable when they are loaded, so it is
# overridden here, before numpy is imported (the N>1 reference arm ran on one thread).
if any(a == "reference" or a.endswith("=reference") for a in sys.argv[1:]):
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(_host_cores())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ccsd_t_fp64_tflops"
UNIT = "TFLOP/s"


def t_flops_per_triple(o, v):
    return 2.0 * v * (v * (v - 1) // 2) * 3 * (v + o)


def make_config(o, v, world):
    nt = o * (o - 1) * (o - 2) // 6
    return {"workload": f"(T) energy, synthetic closed-shell o={o} v={v} (BASELINE.json configs[4]), antisymmetrised "
                        "counter-based inputs, i<j<k triples dealt round-robin over ranks",
            "o": o, "v": v, "triples_total": nt, "alg_flops_per_triple": t_flops_per_triple(o, v),
            "parallelism": f"triples x{world}",
            "l2": "each triple streams 0.84 GB of B rows (>> 126 MB L2); no flush needed"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--o", type=int, default=40)
    ap.add_argument("--v", type=int, default=400)
    ap.add_argument("--triples-per-rank", type=int, default=247)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ccsd", action="store_true")
    ap.add_argument("--no-peak", action="store_true", help="skip the cuBLAS FP64 yardstick (profiler runs)")
    ap.add_argument("--cpu-triples", type=int, default=None, help="triples per CPU sample (default: 3 for cpu_baseline, "
                    "1 per step for --impl reference)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed region (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])), (mx := max(mx, float(r[1])))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------
# CPU path (the oracle; only cpu_baseline and --impl reference may touch it)
# ----------------------------------------------------------------------------------------
CPU_OCC = (0, 1, 2, 20, 39)  # CPU samples draw their triples from these occupied indices


class CpuTriples:
    """Per-triple blocked CPU restatement (numpy/OpenBLAS, all host threads) on sampled i<j<k
    triples of the same synthetic workload.  Only the tensor slabs those triples read are
    generated (same counter-based values as the full tensors); generation is not timed."""

    def __init__(self, o, v):
        from oracle import np_oracle as npo
        self.npo, self.o, self.v = npo, o, v
        occ = sorted({min(x, o - 1) for x in CPU_OCC})
        self.triples = [t for t in npo.sorted_triples(o).tolist() if set(t) <= set(occ)]
        s = npo.synthetic_t_inputs(o, v, occ=occ)
        self.packed = npo.PackedTriples(*(s[k] for k in ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")))
        self.packed.prepare(occ)  # one-off b<c packing of the slabs (amortised over ~740 triples each in the full job)
        self.cursor = 0

    def run(self, ntrip):
        picked, e = [], 0.0
        t0 = time.perf_counter()
        for _ in range(ntrip):
            i, j, k = self.triples[self.cursor % len(self.triples)]
            self.cursor += 1
            e += self.packed.triple(i, j, k)
            picked.append((i, j, k))
        dt = time.perf_counter() - t0
        return dt, ntrip * t_flops_per_triple(self.o, self.v), picked, e


def host_threads():
    """Threads the BLAS behind numpy actually uses (raised to every host core if something capped it)."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        if max((p.get("num_threads", 1) for p in threadpool_info()), default=1) < _host_cores():
            threadpool_limits(_host_cores())  # stays in force for the rest of the process
        n = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
        return int(n)
    except Exception:
        return _host_cores()


def run_reference(args, rank):
    args.cpu_triples = args.cpu_triples or 1
    """--impl reference: the reference's CPU algorithm for this path, on the host cores.
    The reference's own CCSD_T (compiled from its unmodified sources into oracle/_ref, see
    oracle/ref.py) cannot run ANY sample of this workload: it materialises four o^3 v^3 tensors
    (131 TB at o=40/v=400) and its hole term sums over every occupied m, so no sub-block of the
    occupied space is a sample of the job.  The timed path is therefore the oracle port
    (oracle/np_oracle.py: per-triple GEMM on OpenBLAS with every host core -- the same arithmetic
    restated so that it fits in memory, and a much faster CPU code than the reference's own).
    For the record the line also carries `reference_own_code`: oracle/_ref's CCSD_T timed at a
    shape it can hold (o=8, v=40), quoted in the same algorithmic-flop metric."""
    if rank != 0:
        return
    o, v = args.o, args.v
    cores = host_threads()
    cpu = CpuTriples(o, v)
    for _ in range(min(args.warmup, 1)):  # one warm-up triple is enough to spin up the BLAS threads
        cpu.run(1)
    tot_t, flops = 0.0, 0.0
    for _ in range(args.steps):
        dt, fl, _, _ = cpu.run(args.cpu_triples)
        tot_t += dt
        flops += fl
    val = flops / tot_t / 1e12
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(o, v, args.gpus), "arm": {"triples_per_step": args.cpu_triples},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.cpu_triples} sampled i<j<k triples per step x {args.steps} steps of the o={o}/v={v} "
                                       "workload; per-triple M=v,N=v(v-1)/2,K=3(v+o) six DGEMMs on numpy/OpenBLAS from pre-packed slabs + OpenMP a<b<c energy sum"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    own = reference_own_code_sample()
    if own:
        line["reference_own_code"] = own
    emit(line)


def reference_own_code_sample(o=8, v=40):
    """The reference's own CCSD_T (cc.cpp:454-560, unmodified source in oracle/_ref/libqcp2_ref.so) on a
    synthetic input of the same family at a shape whose four o^3 v^3 tensors fit in host memory."""
    try:
        from oracle import np_oracle as npo
        from oracle import ref
        if not ref.available():
            return None
        s = npo.synthetic_t_inputs(o, v)
        t0 = time.perf_counter()
        e = ref.ccsd_t(s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], np.diag(s["eps_o"]).copy(), np.diag(s["eps_v"]).copy())
        dt = time.perf_counter() - t0
        e_port = npo.ccsd_t_blocked(s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"], True)
        fl = math.comb(o, 3) * t_flops_per_triple(o, v)
        return {"kind": "reference", "o": o, "v": v, "seconds": dt, "value": fl / dt / 1e12, "unit": UNIT, "e_t": e,
                "abs_diff_vs_port": abs(e - e_port),
                "note": "reference's own CCSD_T source; its btas::contract runs on the shim's OpenMP GEMM (BTAS/BLAS absent), "
                        "so this understates a BLAS-linked build; executes 18 o^3 v^4 + 18 o^4 v^3 flops for the algorithmic "
                        "C(o,3) 2 v [v(v-1)/2] 3(v+o)"}
    except Exception as exc:  # the reference arm must never fail because of this extra
        return {"kind": "reference", "error": str(exc)}


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
T_GEMM_PROFILES = ("profiles/r2_t_gemm_tma_ncu_full.txt", "profiles/r1_t_gemm_tma_ncu_full.txt")
E2E_N1_RECORDS = ("profiles/r2_bench_n1.json", "profiles/r1_bench_n1_final.json")


def tracked_gemm_traffic():
    """DRAM bytes per triple of the (T) GEMM, read from the newest tracked `ncu --set full` summary
    (tools/ncu_summary.py output): dram__bytes_read.sum + dram__bytes_write.sum of one launch divided by the
    triples of that launch (grid.z).  Returns (bytes_per_triple, file) or (None, None)."""
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for rel in T_GEMM_PROFILES:
        try:
            txt = open(os.path.join(ROOT, rel)).read()
        except OSError:
            continue
        for block in txt.split("kernel:")[1:]:
            if "t_gemm_tma_kernel" not in block.splitlines()[0]:
                continue
            grid = re.search(r"grid/block:\s*\((\d+),\s*(\d+),\s*(\d+)\)", block)
            rd = re.search(r"dram__bytes_read\.sum\s+([0-9.]+)\s+(\w+)", block)
            wr = re.search(r"dram__bytes_write\.sum\s+([0-9.]+)\s+(\w+)", block)
            if grid and rd and wr:
                total = float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)]
                return total / max(int(grid.group(3)), 1), rel
    return None, None


def committed_n1_e2e():
    """E(T) and wall time of the whole job on ONE GPU from the newest tracked record (the bit-identity and
    strong-scaling yardsticks for the N>1 lines)."""
    for rel in E2E_N1_RECORDS:
        try:
            rec = json.loads(open(os.path.join(ROOT, rel)).read().strip().splitlines()[-1])
            if rec.get("n_gpus") == 1 and rec.get("e2e"):
                return rec["e2e"]["e_t"], rec["e2e"]["seconds"], rec["config"]["o"], rec["config"]["v"], rel
        except (OSError, ValueError, KeyError, IndexError):
            continue
    return None, None, None, None, None


def cublas_fp64_peak(torch, n=8192):
    """FP64 roofline denominator measured live: cuBLAS DGEMM via torch.matmul (yardstick only)."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")

    def loop(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / iters

    loop(2)
    burst = min(loop(1) for _ in range(5))
    sustained = loop(60)  # ~2 s back to back
    fl = 2.0 * n ** 3
    del a, b, c
    return fl / burst / 1e12, fl / sustained / 1e12


LAUNCH_LISTS = ("profiles/r2_launches_ccsd.csv",)


def ccsd_roofline(torch, qcp2_b200, ctx, so, foo, fov, fvv, guess, T1, T2, o, v, s_per_iter, peak_tflops):
    """Roofline of one CCSD iteration of the formulation actually RUN (pair-packed ladder family, factorised t1.t1
    terms): the library records every product of one iteration (qcp2_gemm_log); from that record
      executed flops = sum 2 M N K, operand bytes = sum 8 (M K + K N + M N (1 + [beta != 0])) (each operand once),
      bound = sum over products of max(flops / FP64 peak, bytes / HBM) -- the time a perfect kernel per product needs.
    The dominant kernel (the packed tau product against [<ab||ef>; <am||ef>; <mn||ef>]) is then timed alone, live,
    with CUDA events at its recorded shape, and the per-kernel shares of a tracked ncu launch list are attached."""
    import ctypes as C
    p = qcp2_b200._p
    n2 = o + v
    E, it = C.c_double(), C.c_int()
    ctx.gemm_log(True)
    ctx.check(ctx.lib.qcp2_ccsd(ctx.h, p(so), n2, None, p(foo), p(fov), p(fvv), p(guess), o, v, 2, C.c_double(0.0),
                                p(T1), p(T2), C.byref(E), C.byref(it), None, None))
    recs = ctx.gemm_log_read()
    ctx.gemm_log(False)
    hbm = 6459.6e9
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
        hbm_src = "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        hbm_src = "round-1 MEASURED_PEAKS.json value (file absent on this box)"
    peak = (peak_tflops or 35.4) * 1e12
    fl = by = bound = 0.0
    dom = None
    for r in recs:
        f = 2.0 * r["M"] * r["N"] * r["K"] * r["batch"]
        b = 8.0 * (r["M"] * r["K"] + r["K"] * r["N"] + r["M"] * r["N"] * (2 if r["beta"] else 1)) * r["batch"]
        fl, by, bound = fl + f, by + b, bound + max(f / peak, b / hbm)
        if dom is None or f > dom[0]:
            dom = (f, r)
    res = {"executed_gflop_per_iter": fl / 1e9, "operand_gbytes_per_iter": by / 1e9, "products_per_iter": len(recs),
           "bound_s_per_iter": bound, "frac_of_bound": bound / s_per_iter if s_per_iter else None,
           "executed_tflops": fl / s_per_iter / 1e12 if s_per_iter else None,
           "peaks": {"fp64_tflops": peak / 1e12, "fp64_source": "cuBLAS DGEMM 8192^3 measured in this run (sustained)" if peak_tflops else "35.4 (round-1 yardstick; --no-peak)",
                     "hbm_gbs": hbm / 1e9, "hbm_source": hbm_src},
           "note": "bound counts the GEMMs only (each operand streamed once at the measured HBM rate, flops at the measured cuBLAS FP64 "
                   "rate); the ~25 elementwise / packing kernels of an iteration are not in it"}
    if dom:
        f, r = dom
        M, N, K = r["M"], r["N"], r["K"]
        ld = K + (K & 1)
        fdev = dict(dtype=torch.float64, device="cuda")
        A = torch.randn((M, ld), **fdev)
        B = torch.randn((N, ld), **fdev)
        Cm = torch.empty((M, N + (N & 1)), **fdev)
        torch.cuda.current_stream().synchronize()
        ctx.dgemm_ex(0, 1, M, N, K, 1.0, A, ld, B, ld, 0.0, Cm, N + (N & 1), reps=2)
        t = ctx.dgemm_ex(0, 1, M, N, K, 1.0, A, ld, B, ld, 0.0, Cm, N + (N & 1), reps=10)
        res["dominant_kernel"] = {"kernel": "ws_dgemm_kernel<WsShape<1,8,6,4>,A [M][K],B [N][K]> (48x256 CTA tile, 5-stage TMA/mbarrier ring, "
                                            "8 DMMA.8x8x4 consumer warps, in-kernel ordered split-K)",
                                  "what": "packed tau [i<j][e<f] x [<ab||ef>; <am||ef>; <mn||ef>]^T", "M": M, "N": N, "K": K,
                                  "tile": [r["tile_m"], r["tile_n"]], "split_k": r["split"], "avg_us": t * 1e6, "tflops": f / t / 1e12,
                                  "frac": f / t / peak, "bound": "tensor",
                                  "alg_flops_per_launch": f, "timed": "alone, 10 launches, CUDA events on the launching stream"}
        del A, B, Cm
    for rel in LAUNCH_LISTS:
        path = os.path.join(ROOT, rel)
        if os.path.exists(path):
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import launch_table
            seq = launch_table.load(path)
            idx = [i for i, (n, _, _) in enumerate(seq) if n.startswith("ccsd_energy")]
            one = seq[idx[-2]:idx[-1]] if len(idx) >= 2 else seq
            tot = sum(t for _, t, _ in one)
            agg = {}
            for n, t, _ in one:
                agg[n] = agg.get(n, 0.0) + t
            res["launch_list"] = {"file": rel, "launches": len(one), "serialised_us": tot,
                                  "top": [{"kernel": n, "us": t, "share": t / tot} for n, t in sorted(agg.items(), key=lambda x: -x[1])[:8]]}
            break
    return res


def ccsd_seconds_per_iter(torch, qcp2_b200, ctx, o=10, v=162, iters=8, cpu_sample=False, sharded=False, roofline=False,
                          peak_tflops=None):
    """Secondary metric: CCSD seconds/iteration on the CH4/cc-pVTZ shape (o=10, v=162).  The integrals are a
    random-init spin-orbital tensor <pq||rs> with the permutational symmetries every real input has
    (antisymmetric in pq and in rs, symmetric under pq<->rs); it is sliced on the device and iterated through
    the C ABI in device pointer mode (qcp2_ccsd), a fixed number of iterations (conv = 0)."""
    import ctypes as C
    n2 = o + v
    g = torch.Generator(device="cuda").manual_seed(7)
    f = dict(dtype=torch.float64, device="cuda")
    so = torch.randn((n2,) * 4, generator=g, **f) * 0.001  # small enough for the Jacobi iteration to stay contractive
    so = so - so.transpose(0, 1)
    so = so - so.transpose(2, 3)
    so = (so + so.permute(2, 3, 0, 1)).contiguous()
    if sharded:  # collective run: every rank must hold the very same tensor
        ctx.comm_broadcast(so, root=0)
        ctx.set_ccsd_sharding(True)
    eo = torch.linspace(-2.0, -0.5, o, **f)
    ev = torch.linspace(0.5, 3.0, v, **f)
    foo, fvv, fov = torch.diag(eo).contiguous(), torch.diag(ev).contiguous(), torch.zeros((o, v), **f)
    D = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
    guess = (so[:o, :o, o:, o:] / D).contiguous()
    T1, T2 = torch.empty((o, v), **f), torch.empty((o, o, v, v), **f)
    E, it = C.c_double(), C.c_int()
    torch.cuda.current_stream().synchronize()
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    p = qcp2_b200._p
    launches0 = ctx.kernel_launches
    secs = []
    for rep in range(2):  # first call warms the workspace pool
        ctx.check(ctx.lib.qcp2_ccsd(ctx.h, p(so), n2, None, p(foo), p(fov), p(fvv), p(guess), o, v, iters, C.c_double(0.0),
                                    p(T1), p(T2), C.byref(E), C.byref(it), None, None))
        secs.append(float(ctx.lib.qcp2_ccsd_last_loop_seconds(ctx.h)) / iters)
    flops = 4.0 * o * o * v ** 4 + 4.0 * o * v ** 4 + 20.0 * o ** 3 * v ** 3 + 4.0 * o ** 4 * v * v + 18.0 * o * o * v ** 3
    out = {"s_per_iter": secs[-1], "o": o, "v": v, "iters_timed": iters, "tflops_reference_formulation": flops / secs[-1] / 1e12,
           "data": "random-init antisymmetrised <pq||rs>, CH4/cc-pVTZ shape", "launches_per_iter": (ctx.kernel_launches - launches0) // (2 * iters),
           "e_after_iters": E.value if math.isfinite(E.value) else None}
    if roofline and not sharded:
        out["roofline"] = ccsd_roofline(torch, qcp2_b200, ctx, so, foo, fov, fvv, guess, T1, T2, o, v, secs[-1], peak_tflops)
    if cpu_sample:
        # one iteration of the numpy/einsum CPU port (BLAS on all host cores) on the same blocks
        from oracle import np_oracle as npo
        sl = lambda c: slice(0, o) if c == "o" else slice(o, n2)
        I = {k: so[sl(k[0]), sl(k[1]), sl(k[2]), sl(k[3])].contiguous().cpu().numpy() for k in qcp2_b200.INT_NAMES}
        foo_h, fov_h, fvv_h, t2_h = foo.cpu().numpy(), fov.cpu().numpy(), fvv.cpu().numpy(), guess.cpu().numpy()
        t1_h = np.zeros((o, v))
        t0 = time.perf_counter()
        X = npo.intermediates(t1_h, t2_h, I, foo_h, fov_h, fvv_h)
        t1n = npo.new_T1(t1_h, t2_h, I, X, fov_h, eo.cpu().numpy(), ev.cpu().numpy())
        npo.new_T2(t1n, t2_h, I, X, eo.cpu().numpy(), ev.cpu().numpy())
        out["cpu_s_per_iter"] = time.perf_counter() - t0
        out["cpu_kind"] = "port (oracle/np_oracle.py, einsum->OpenBLAS), %d threads, 1 iteration" % host_threads()
    if sharded:
        ctx.set_ccsd_sharding(False)
    del so
    return out


class SharedHostTensors:
    """The job's host tensors as ONE copy in POSIX shared memory (/dev/shm), mapped by every rank of the node and pinned in each
    process with cudaHostRegister -- what a multi-process host program does so that every GPU can pull its share of the
    inputs over its own PCIe link (qcp2_ccsd_t_mgpu with the tensors passed on every rank).  Set-up is outside the timed region,
    like the pinned allocation of the single-process case.  `ok` is the same on every rank."""

    def __init__(self, torch, dist, rank, dev_tensors):
        self.torch, self.dist, self.rank = torch, dist, rank
        self.arrays, self.paths, self.registered = {}, [], []
        tag = os.environ.get("MASTER_PORT", "0") + "_" + str(os.getppid())
        good = 1.0
        try:
            for k, t in dev_tensors.items():
                path = f"/dev/shm/qcp2_bench_{tag}_{k}.bin"
                self.paths.append(path)
                if rank == 0:
                    m = np.memmap(path, dtype=np.float64, mode="w+", shape=(t.numel(),))
                    torch.from_numpy(m).copy_(t.reshape(-1))
                    self.arrays[k] = m.reshape(tuple(t.shape))
        except Exception as ex:  # no /dev/shm, not enough room ...
            sys.stderr.write(f"[bench] shared host tensors unavailable: {ex}\n")
            good = 0.0
        dist.barrier()
        try:
            if rank != 0 and good:
                for k, t in dev_tensors.items():
                    m = np.memmap(f"/dev/shm/qcp2_bench_{tag}_{k}.bin", dtype=np.float64, mode="r+", shape=(t.numel(),))
                    self.arrays[k] = m.reshape(tuple(t.shape))
            if good:
                rt = torch.cuda.cudart()
                for k, a in self.arrays.items():
                    if a.nbytes >= (1 << 20):  # pin the large ones; the rest is staged by the driver
                        rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
                        if int(rc) != 0:
                            raise RuntimeError(f"cudaHostRegister({k}) -> {rc}")
                        self.registered.append(a.ctypes.data)
        except Exception as ex:
            sys.stderr.write(f"[bench] rank {rank}: shared host tensors unavailable: {ex}\n")
            good = 0.0
        flag = torch.tensor([good], dtype=torch.float64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = bool(flag.item() > 0.5)

    def close(self):
        rt = self.torch.cuda.cudart()
        for p in self.registered:
            rt.cudaHostUnregister(p)
        self.registered = []
        self.arrays = {}
        self.dist.barrier()
        if self.rank == 0:
            for path in self.paths:
                try:
                    os.unlink(path)
                except OSError:
                    pass
        self.paths = []


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout: anything libraries print there (NCCL's version banner,
    BLAS chatter) is sent to stderr for the duration of the run; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import qcp2_b200
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    o, v, B = args.o, args.v, args.triples_per_rank
    ctx = qcp2_b200.Context(local)
    # one dedicated stream shared by torch (NCCL collectives order against it), the library and
    # the timing events; torch's default stream has handle 0, which the library reads as "own stream"
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    if world > 1:
        # the library's own NCCL communicator (qcp2_comm_*); torch.distributed only carries the
        # 128-byte bootstrap id, the barriers and the max-over-ranks of the timings
        ctx.comm_init_from_torch(dist)
        # communicator warm-up (NCCL builds its channels lazily, per collective, on first use): one tiny broadcast from every
        # root, one all-gather, one all-reduce -- so that no connection set-up lands inside a timed region
        wbuf = torch.zeros(8 * world, dtype=torch.float64, device="cuda")
        for r in range(world):
            ctx.comm_broadcast(wbuf, root=r)
        ctx.comm_allgather(wbuf[8 * rank:8 * rank + 8], wbuf)
        ctx.comm_allreduce_sum(wbuf)
        ctx.synchronize()

    # ---- inputs: rank 0 generates, NCCL broadcast replicates (outside the timed region for `value`)
    d = qcp2_b200.synth_t_inputs_device(ctx, o, v)
    names = ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")
    if world > 1:
        if rank != 0:
            for k in names:
                d[k].zero_()
        for k in names:
            ctx.comm_broadcast(d[k], root=0)
    plan = qcp2_b200.DeviceTriples(ctx, *(d[k] for k in names), qcp2_b200.T_SYMMETRIC)
    nt = ctx.t_count(o, qcp2_b200.T_SYMMETRIC)
    e_trip = torch.zeros(nt, dtype=torch.float64, device="cuda")
    fpt = t_flops_per_triple(o, v)

    def step(s):
        first = (s * B * world) % max(nt, 1) + rank
        e_trip.zero_()
        plan.run(first, world, B, e_trip)
        tm = plan.timings()
        if world > 1:
            ctx.comm_allreduce_sum(e_trip)  # one slot per triple, written by exactly one rank: exact
        total = qcp2_b200.shard.reduce_triple_energies(e_trip, None)  # fixed-order host sum
        done = len(range(first, nt, world)[:B])
        return total, tm, done

    for s in range(args.warmup):
        step(s)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gemm_s, gemm_n, done_local, energy_acc = 0.0, 0, 0, 0.0
    for s in range(args.steps):
        total, tm, done = step(args.warmup + s)
        gemm_s += tm["gemm_seconds"]
        gemm_n += tm["gemm_launches"]
        done_local += done
        energy_acc += total
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed = torch.tensor([e0.elapsed_time(e1) * 1e-3, float(done_local), gemm_s, float(gemm_n)], dtype=torch.float64,
                           device="cuda")
    if world > 1:
        mx = elapsed.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = elapsed.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_max, done_all = float(mx[0]), float(sm[1])
    else:
        t_max, done_all = float(elapsed[0]), float(elapsed[1])
    value = done_all * fpt / t_max / 1e12

    # ---- roofline of the dominant kernel (the segmented-K DMMA GEMM), this rank
    peak_burst, peak_sus = (0.0, 0.0) if args.no_peak else cublas_fp64_peak(torch)
    achieved = done_local * fpt / gemm_s / 1e12 if gemm_s > 0 else 0.0
    # DRAM bytes per triple of the GEMM, read from the tracked ncu --set full summary (captured at o=40, v=400);
    # algorithmic bytes per triple: B rows 843 MB + X 255 MB + A 5 MB = 1.10 GB
    dram_per_triple, traffic_file = tracked_gemm_traffic() if (o, v) == (40, 400) else (None, None)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_sus, "unit": "TFLOP/s",
                "frac": achieved / peak_sus if peak_sus else None,
                "traffic": dram_per_triple * done_local / max(gemm_n, 1) if dram_per_triple else None,
                "traffic_source": traffic_file,
                "kernel": "t_gemm_tma_kernel (80x256x16 CTA tile, 4-stage cp.async.bulk/mbarrier ring, 8 DMMA.8x8x4 consumer warps)",
                "avg_launch_ms": 1e3 * gemm_s / max(gemm_n, 1),
                "launches": gemm_n, "alg_flops_per_launch": done_local * fpt / max(gemm_n, 1),
                "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul measured in this run (sustained ~2 s loop; burst %.2f); "
                               "MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2" % peak_burst}

    # ---- e2e: the whole (T) job through the C ABI with HOST (pinned) buffers
    e2e = None
    if not args.no_e2e:
        host, shared = {}, None
        if world > 1:
            shared = SharedHostTensors(torch, dist, rank, {k: d[k] for k in names})
            if shared.ok:
                host = shared.arrays
            else:
                shared.close()
                shared = None
        if shared is None and rank == 0:
            for k in names:
                h = torch.empty(d[k].shape, dtype=torch.float64, pin_memory=True)
                h.copy_(d[k])
                host[k] = h.numpy()
        plan.close()
        del plan
        d_numel = {k: d[k].numel() for k in names}
        h2d = sum(d_numel.values()) * 8
        for k in names:
            d[k] = None
        torch.cuda.empty_cache()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world == 1:
            ctx.set_pointer_mode(qcp2_b200.HOST)
            e_job = ctx.CCSD_T(*(host[k] for k in names), qcp2_b200.T_AUTO)
            ctx.set_pointer_mode(qcp2_b200.DEVICE)
        else:
            # qcp2_ccsd_t_mgpu: rank 0 stages its pinned host tensors, ncclBroadcast replicates them,
            # every rank evaluates its round-robin share, one ncclAllReduce of the per-triple vector
            ctx.set_pointer_mode(qcp2_b200.HOST)
            # every rank maps the ONE shared-memory host copy (pinned in each process): the library then deals the <vo||vv>
            # slabs over the ranks' own PCIe links; without shared memory only rank 0 holds (and uploads) the tensors
            src = [host[k] if (rank == 0 or shared is not None) else None for k in names]
            e_job = ctx.CCSD_T_mgpu(*src, o, v, qcp2_b200.T_AUTO, root=0)
            ctx.set_pointer_mode(qcp2_b200.DEVICE)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ref_e, ref_s, ref_o, ref_v, ref_file = committed_n1_e2e()
        comparable = ref_e is not None and (ref_o, ref_v) == (o, v)
        if shared is not None:
            # every large tensor crosses PCIe once in total (1/W of it per rank); T1 and the orbital energies on every rank
            h2d = sum(d_numel.values()) * 8 + (world - 1) * (d_numel["T1"] + d_numel["eps_o"] + d_numel["eps_v"]) * 8
            shared.close()
        e2e = {"value": nt * fpt / float(dt[0]) / 1e12, "unit": UNIT,
               # E(T) of the whole job against the committed one-GPU value, bit for bit (fixed triple->slot map + fixed-order sum)
               "e_t_matches_n1": (e_job == ref_e) if comparable else None, "e_t_n1": ref_e if comparable else None,
               # whole-job strong scaling: T(1 GPU, committed record) / (N x T(N GPUs, this run)); at N=1 it compares boxes
               "strong_scaling_efficiency": ref_s / (world * float(dt[0])) if comparable else None,
               "n1_record": ref_file if comparable else None, "h2d_bytes_per_step": h2d,
               "host_buffers": ("one POSIX shared-memory copy mapped and pinned (cudaHostRegister) by every rank; <vo||vv> slabs dealt over the "
                                "ranks' PCIe links" if shared is not None else "pinned on rank 0" if world > 1 else "pinned"),
               "d2h_bytes_per_step": 8 * (nt + 1), "steps": 1, "seconds": float(dt[0]), "e_t": e_job,
               "what": "whole (T) job (all %d triples) via qcp2_ccsd_t with pinned host buffers" % nt if world == 1 else
                       ("whole (T) job via qcp2_ccsd_t_mgpu: every rank uploads the small tensors and its share of the <vo||vv> slabs from the "
                        "shared pinned host copy -> per-slab ncclBroadcast from the uploading rank -> round-robin triples -> ncclAllReduce"
                        if shared is not None else
                        "whole (T) job via qcp2_ccsd_t_mgpu: rank-0 pinned host buffers -> H2D -> ncclBroadcast -> round-robin triples -> ncclAllReduce")}

    ccsd = None
    if not args.no_ccsd and (rank == 0 or world > 1):
        try:  # N > 1: collective (row/column-sharded products of every iteration over the library's communicator)
            ccsd = ccsd_seconds_per_iter(torch, qcp2_b200, ctx, cpu_sample=(world == 1 and not args.no_cpu_baseline), sharded=(world > 1),
                                         roofline=(world == 1), peak_tflops=peak_sus or None)
            ccsd["parallelism"] = ("single GPU" if world == 1 else
                                   "large products of each iteration split over %d ranks along their longer free dimension, one "
                                   "ncclAllGather each (qcp2_set_ccsd_sharding); everything else replicated" % world)
        except Exception as ex:  # never lose the headline line to the secondary metric
            ccsd = {"error": str(ex)[:200]}

    cpu = None
    args.cpu_triples = args.cpu_triples or 3
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dt, fl, pick, _ = CpuTriples(o, v).run(args.cpu_triples)
        cpu = {"value": fl / dt / 1e12, "unit": UNIT, "cores": host_threads(), "kind": "port",
               "sample": f"{len(pick)} sampled i<j<k triples {pick} of the same o={o}/v={v} workload, per-triple DGEMM on "
                         f"numpy/OpenBLAS + OpenMP a<b<c energy sum, {dt:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_max / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": make_config(o, v, world), "arm": {"triples_per_rank_per_step": B},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "ccsd": ccsd, "energy_checksum": energy_acc}
        emit(line)
    ctx.close()  # also destroys the library's communicator
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
