"""Runs BASELINE.json configs 1-4 end to end (host SCF -> GPU MP2/CCSD/(T) through the C ABI)
next to the CPU oracles on the same boundary inputs and writes gpurun_out/config_parity.json:
energies, |GPU - CPU|, CCSD iterations, device seconds per iteration, CPU seconds.
Config 3 (CH4/cc-pVTZ) uses the numpy/einsum oracle (valid for RHF inputs; the literal C++ oracle
needs ~10 min per few iterations there) and scf_conv = 1e-10 (1e-12 is below the double-precision
noise floor of ||dD|| at n = 86)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qcp2_b200  # noqa: E402
from helpers import fock_blocks, molecule_case, oracle_pipeline  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CONFIGS = [
    ("1 water/STO-3G RHF", dict(xyz="water.xyz", basis="sto-3g"), "cpp"),
    ("2 water/cc-pVDZ RHF", dict(xyz="water.xyz", basis="cc-pvdz"), "cpp"),
    ("4 NH2/cc-pVDZ UHF", dict(xyz="nh2.xyz", basis="cc-pvdz", uhf=True, multiplicity=2), "cpp"),
    ("3 CH4/cc-pVTZ RHF", dict(xyz="methane.xyz", basis="cc-pvtz", conv=1e-10, maxiter=400), "numpy"),
]


def main():
    only = sys.argv[1:]
    ctx = qcp2_b200.Context(0)
    out = {}
    for name, cfg, cpu_kind in CONFIGS:
        if only and not any(name.startswith(x) for x in only):
            continue
        t0 = time.time()
        case = molecule_case(**cfg)
        t_scf = time.time() - t0
        uhf = cfg.get("uhf", False)
        no, nv = case["no"], case["nv"]
        ea, eb = (case["eps_a"], case["eps_b"]) if uhf else (case["eps"], case["eps"])
        foo, fov, fvv = fock_blocks(ea, eb, no)
        eo, ev = np.diag(foo).copy(), np.diag(fvv).copy()
        # ---- GPU
        t0 = time.time()
        e_mp2, so, T2g = ctx.MP2(case["ao"], case["Ca"] if uhf else case["C"], case.get("Cb"), ea, eb, no, nv, uhf=uhf)
        t_mp2 = time.time() - t0
        t0 = time.time()
        res = ctx.CCSD(foo, None if uhf else fov, fvv, T2g, 200, 1e-12, so_ints=so, want_sliced=True)
        t_cc = time.time() - t0
        S = res["sliced_ints"]
        t0 = time.time()
        mode = ctx.t_select_mode(res["T2"], S["oovv"], S["vovv"], S["ovoo"])
        e_t = ctx.CCSD_T(res["T1"], res["T2"], S["oovv"], S["vovv"], S["ovoo"], eo, ev)
        t_t = time.time() - t0
        t0 = time.time()
        fused = ctx.post_hf(case["ao"], case["Ca"] if uhf else case["C"], case.get("Cb"), ea, eb if uhf else None, no, nv,
                            uhf=uhf, stages=3, maxiter=200, cc_conv=1e-12)
        t_fused = time.time() - t0
        rec = dict(n=case["n"], o=no, v=nv, e_scf=case["e_scf"], scf_seconds=t_scf,
                   fused=dict(call_seconds=t_fused, e_mp2=fused["e_mp2"], e_ccsd=fused["e_ccsd"], e_t=fused["e_t"],
                              iters=fused["iters"]),
                   gpu=dict(e_mp2=e_mp2, e_ccsd=res["ccsd_energy"], e_t=e_t, iters=res["iters"],
                            ccsd_loop_seconds=res["loop_seconds"], ccsd_s_per_iter=res["loop_seconds"] / max(res["iters"], 1),
                            mp2_call_seconds=t_mp2, ccsd_call_seconds=t_cc, t_call_seconds=t_t,
                            t_mode="literal" if mode == qcp2_b200.T_LITERAL else "symmetric"))
        print(name, json.dumps(rec["gpu"]), flush=True)
        # ---- CPU oracle on the same boundary inputs
        t0 = time.time()
        if cpu_kind == "cpp":
            ref = oracle_pipeline(case, maxiter=200, conv=1e-12, with_t=False)
            I = ref["ints"]
            e_t_ref = orc.ccsd_t_blocked(ref["T1"], ref["T2"], I["oovv"], I["vovv"], I["ovoo"], eo, ev, npo.all_triples(no))
            cpu = dict(kind="oracle.cpp (literal)", e_mp2=ref["e_mp2"], e_ccsd=ref["e_ccsd"], e_t=e_t_ref, iters=ref["iters"],
                       hist_maxdiff=float(np.max(np.abs(np.array(ref["e_hist"]) - res["e_hist"]))))
        else:
            # ~15 min of CPU: use the committed CPU-oracle numbers (tests/golden/make_ch4_golden.py)
            g = json.load(open(os.path.join(ROOT, "tests", "golden", "ch4_ccpvtz_oracle.json")))
            hist = g["e_hist"]
            cpu = dict(kind="np_oracle (einsum), committed fixture tests/golden/ch4_ccpvtz_oracle.json", e_mp2=g["e_mp2"],
                       e_ccsd=g["e_ccsd"], e_t=g["e_t"], iters=g["iters"], fixture_seconds=g["seconds"],
                       hist_maxdiff=float(np.max(np.abs(np.array(hist) - res["e_hist"]))) if len(hist) == len(res["e_hist"]) else None)
        cpu["seconds"] = time.time() - t0
        rec["cpu"] = cpu
        rec["abs_diff"] = dict(e_mp2=abs(e_mp2 - cpu["e_mp2"]), e_ccsd=abs(res["ccsd_energy"] - cpu["e_ccsd"]),
                               e_t=abs(e_t - cpu["e_t"]))
        rec["parity_1e-10"] = all(v < 1e-10 for v in rec["abs_diff"].values()) and cpu["iters"] == res["iters"]
        print(name, "CPU", json.dumps(cpu), "diff", json.dumps(rec["abs_diff"]), "PARITY" if rec["parity_1e-10"] else "MISMATCH", flush=True)
        out[name] = rec
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_parity.json"), "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
