"""Roofline evidence for the bandwidth-bound kernels and the quarter transformation
(SURVEY.md section 8d): achieved GB/s = 8 B x (elements read once + elements written once) / CUDA-event
time, against the measured HBM copy bandwidth (MEASURED_PEAKS.json, 6459.6 GB/s on this pool) --
and algorithmic TFLOP/s (8 n^5) for transform_ao_mo.  Everything runs through the C ABI in device
pointer mode on the CH4/cc-pVTZ shape (n = 86, o = 10, v = 162).  Writes gpurun_out/bandwidth_kernels.json."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qcp2_b200  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


ONCE = os.environ.get("QCP2_BW_ONCE") == "1"  # under ncu: one launch of each kernel is enough


def timed(fn, reps=5):
    fn()
    if ONCE:
        return 1.0
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    ctx = qcp2_b200.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    p, lib, h = qcp2_b200._p, ctx.lib, ctx.h
    f = dict(dtype=torch.float64, device="cuda")
    n, o, v = 86, 10, 162
    n2 = 2 * n
    rows = []

    def rec(name, seconds, nbytes=None, flops=None):
        r = dict(kernel=name, ms=seconds * 1e3)
        if nbytes is not None:
            r["GB"] = nbytes / 1e9
            r["GBps"] = nbytes / seconds / 1e9
            r["frac_of_measured_hbm"] = r["GBps"] / HBM
        if flops is not None:
            r["TFLOPs"] = flops / seconds / 1e12
        rows.append(r)
        print(r, flush=True)

    # local ceilings for context: device-to-device copy (read + write) and pure write (memset) of 7 GB
    big0, big1 = torch.empty(7 * 2 ** 27, **f), torch.empty(7 * 2 ** 27, **f)
    rec("ceiling: torch copy_ 7.5 GB -> 7.5 GB", timed(lambda: big1.copy_(big0)), nbytes=16.0 * big0.numel())
    rec("ceiling: torch zero_ 7.5 GB (write only)", timed(lambda: big1.zero_()), nbytes=8.0 * big0.numel())
    del big0, big1
    ao, Cm = torch.randn((n,) * 4, **f), torch.randn((n, n), **f)
    mo = torch.empty((n,) * 4, **f)
    rec("transform_ao_mo n=86 (4 DMMA GEMMs)", timed(lambda: ctx.check(lib.qcp2_transform_ao_mo(h, p(ao), p(Cm), p(Cm), n, p(mo)))),
        flops=8.0 * n ** 5)
    so = torch.empty((n2,) * 4, **f)
    rec("transform_to_so n=86", timed(lambda: ctx.check(lib.qcp2_transform_to_so(h, p(mo), p(mo), p(mo), n, p(so)))),
        nbytes=8.0 * (n2 ** 4 + n ** 4))
    for spec in ("vvvv", "vovv", "oovv"):
        shape = qcp2_b200.int_shape(spec, o, v)
        out = torch.empty(shape, **f)
        rec(f"slice_ints {spec}", timed(lambda: ctx.check(lib.qcp2_slice_ints(h, p(so), n2, o, v, spec.encode(), p(out)))),
            nbytes=16.0 * out.numel())
        if spec == "oovv":
            oovv = out
    del so, ao, mo
    ts, td = torch.randn((o, v), **f) * 0.01, torch.randn((o, o, v, v), **f) * 0.01
    tau = torch.empty_like(td)
    rec("make_tau", timed(lambda: ctx.check(lib.qcp2_make_tau(h, p(ts), p(td), o, v, 0, p(tau)))), nbytes=16.0 * td.numel())
    tt = torch.empty((o, v, o, v), **f)
    rec("multiply_Ts", timed(lambda: ctx.check(lib.qcp2_multiply_Ts(h, p(ts), o, v, p(tt)))), nbytes=8.0 * tt.numel())
    eps = torch.linspace(-2, 3, n, **f)
    F = torch.empty((n2, n2), **f)
    ctx.check(lib.qcp2_make_so_moes(h, p(eps), p(eps), n, p(F)))
    den = torch.empty_like(td)
    rec("make_denom", timed(lambda: ctx.check(lib.qcp2_make_denom(h, p(F), n2, o, v, p(den)))), nbytes=8.0 * den.numel())
    E = C.c_double()
    T = torch.empty_like(td)
    rec("run_mp2 (divide + energy reduction)", timed(lambda: ctx.check(lib.qcp2_run_mp2(h, p(oovv), p(den), o, v, p(T), C.byref(E)))),
        nbytes=24.0 * td.numel())
    fov = torch.zeros((o, v), **f)
    rec("ccsd_energy (reduction)", timed(lambda: ctx.check(lib.qcp2_ccsd_energy(h, p(ts), p(td), p(oovv), p(fov), o, v, C.byref(E)))),
        nbytes=16.0 * td.numel())
    # permutation kernel through contract(): 'amef' -> 'maef' re-layout dominates a K=10 contraction
    big = torch.randn((v, o, v, v), **f)
    outp = torch.empty((v, v, v, v), **f)
    ext = lambda t: (C.c_longlong * t.dim())(*t.shape)
    rec("contract ma,bmef->abef (permute ov^3 + K=10 GEMM writing v^4)",
        timed(lambda: ctx.check(lib.qcp2_contract(h, C.c_double(1.0), p(ts), ext(ts), b"ma", p(big), ext(big), b"bmef", C.c_double(0.0),
                                                   p(outp), ext(outp), b"abef")), reps=3),
        nbytes=8.0 * (outp.numel() + 3 * big.numel()))
    # larger transposes: permute_axpby via a pure re-layout contraction with a 1x1 "matrix"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(measured_hbm_gbs=HBM, shape=dict(n=n, o=o, v=v), rows=rows),
              open(os.path.join(ROOT, "gpurun_out", "bandwidth_kernels.json"), "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
