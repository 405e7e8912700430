"""GPU parity on the real-molecule configs of BASELINE.json (run with -m gpu): boundary inputs
from the host SCF substrate, then MP2 -> CCSD -> (T) through the C ABI vs the CPU oracle on the
same inputs, 1e-10 Eh.  Config 3 (CH4/cc-pVTZ) takes minutes on the CPU side and lives in
tools/run_configs.py."""
import json
import os
import subprocess

import numpy as np
import pytest
from helpers import ROOT, fock_blocks, molecule_case, oracle_pipeline

import qcp2_b200
from oracle import np_oracle as npo
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
E_TOL = 1e-10

CONFIGS = {
    "water_sto3g_rhf": dict(xyz="water.xyz", basis="sto-3g"),
    "water_ccpvdz_rhf": dict(xyz="water.xyz", basis="cc-pvdz"),
    "nh2_ccpvdz_uhf": dict(xyz="nh2.xyz", basis="cc-pvdz", uhf=True, multiplicity=2),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_config_energy_parity(ctx, name):
    cfg = CONFIGS[name]
    case = molecule_case(**cfg)
    uhf = cfg.get("uhf", False)
    no, nv = case["no"], case["nv"]
    ref = oracle_pipeline(case, maxiter=200, conv=1e-12, with_t=False)
    ea, eb = (case["eps_a"], case["eps_b"]) if uhf else (case["eps"], case["eps"])
    e_mp2, so, T2g = ctx.MP2(case["ao"], case["Ca"] if uhf else case["C"], case.get("Cb"), ea, eb, no, nv, uhf=uhf)
    assert abs(e_mp2 - ref["e_mp2"]) < E_TOL
    foo, fov, fvv = fock_blocks(ea, eb, no)
    res = ctx.CCSD(foo, None if uhf else fov, fvv, T2g, 200, 1e-12, so_ints=so, want_sliced=True)
    assert res["iters"] == ref["iters"]
    assert np.max(np.abs(res["e_hist"] - ref["e_hist"])) < E_TOL
    assert abs(res["ccsd_energy"] - ref["e_ccsd"]) < E_TOL
    S = res["sliced_ints"]
    eo, ev = np.diag(foo).copy(), np.diag(fvv).copy()
    mode = ctx.t_select_mode(res["T2"], S["oovv"], S["vovv"], S["ovoo"])
    assert mode == (qcp2_b200.T_LITERAL if uhf else qcp2_b200.T_SYMMETRIC)
    e_t = ctx.CCSD_T(res["T1"], res["T2"], S["oovv"], S["vovv"], S["ovoo"], eo, ev)
    # CPU (T) on the oracle's own converged amplitudes, per-triple form, no symmetry assumed
    I = ref["ints"]
    e_t_ref = orc.ccsd_t_blocked(ref["T1"], ref["T2"], I["oovv"], I["vovv"], I["ovoo"], eo, ev, npo.all_triples(no))
    assert abs(e_t - e_t_ref) < E_TOL
    if name == "water_sto3g_rhf":  # external anchor (Crawford), see tests/test_known_answers.py
        assert abs(e_mp2 - (-0.049149636120)) < 1e-10 and abs(res["ccsd_energy"] - (-0.070680088376)) < 1e-10
        assert abs(e_t - (-0.000099877272)) < 1e-11


def test_configs_vs_the_reference_itself(ctx):
    """GPU path vs tests/golden/reference_energies.json -- the energies the reference's OWN sources
    (compiled against oracle/ref_shim, see tests/golden/make_reference_golden.py) produce for
    BASELINE.json configs 1, 2 and 4, through the fused qcp2_post_hf entry point.  The UHF config is
    fed the reference's own converged C / eps (reference_nh2_boundary.npz): the as-written UHF path
    depends on the arbitrary signs of the MO vectors, so only identical boundary inputs
    give identical energies.  The AO integrals are bit-identical by construction (same host code)."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_energies.json")))
    for name, cfg in sorted(CONFIGS.items()):
        case = molecule_case(**cfg)
        uhf = cfg.get("uhf", False)
        if uhf:
            b = np.load(os.path.join(ROOT, "tests", "golden", "reference_nh2_boundary.npz"))
            Ca, Cb, ea, eb = b["Ca"], b["Cb"], b["eps_a"], b["eps_b"]
        else:
            assert abs(case["e_scf"] - g[name]["e_scf"]) < 1e-10
            Ca, Cb, ea, eb = case["C"], None, case["eps"], None
        r = ctx.post_hf(case["ao"], Ca, Cb, ea, eb, case["no"], case["nv"], uhf=uhf, stages=3, maxiter=200, cc_conv=1e-12)
        assert abs(r["e_mp2"] - g[name]["e_mp2"]) < E_TOL, name
        assert r["iters"] == g[name]["cc_iters"], name
        assert abs(r["e_ccsd"] - g[name]["e_ccsd"]) < E_TOL, name
        assert abs(r["e_t"] - g[name]["e_t"]) < E_TOL, name
        # RHF inputs carry the spin zero pattern of transform_to_so: pair-packed iteration, one block per spin class (mode 2);
        # the as-written UHF integrals fail the antisymmetry checks and take the dense term list (mode 0)
        assert ctx.ccsd_last_mode() == (0 if uhf else 2), (name, ctx.ccsd_last_mode())


def test_opt_in_diis_reaches_the_same_ccsd_energy_in_fewer_iterations(ctx):
    """DIIS is not in the reference (cc.cpp:426-447 is a plain Jacobi loop) and is OFF by default; switched on it must reach
    the same fixed point (1e-10 Eh) in fewer iterations.  The as-written UHF equations have several fixed points,
    so the dense formulation ignores the request and keeps the reference's trajectory."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_energies.json")))
    for name in ("water_ccpvdz_rhf", "nh2_ccpvdz_uhf"):
        cfg = CONFIGS[name]
        case = molecule_case(**cfg)
        uhf = cfg.get("uhf", False)
        if uhf:
            b = np.load(os.path.join(ROOT, "tests", "golden", "reference_nh2_boundary.npz"))
            Ca, Cb, ea, eb = b["Ca"], b["Cb"], b["eps_a"], b["eps_b"]
        else:
            Ca, Cb, ea, eb = case["C"], None, case["eps"], None
        ctx.set_ccsd_diis(6)
        try:
            r = ctx.post_hf(case["ao"], Ca, Cb, ea, eb, case["no"], case["nv"], uhf=uhf, stages=3, maxiter=200, cc_conv=1e-12)
        finally:
            ctx.set_ccsd_diis(0)
        assert abs(r["e_ccsd"] - g[name]["e_ccsd"]) < E_TOL, (name, r["e_ccsd"], g[name]["e_ccsd"])
        assert abs(r["e_t"] - g[name]["e_t"]) < 1e-9, name
        if uhf:
            assert r["iters"] == g[name]["cc_iters"], (name, r["iters"])
        else:
            assert r["iters"] < g[name]["cc_iters"], (name, r["iters"], g[name]["cc_iters"])
    # and off again: the reference's iteration count is back
    case = molecule_case(**CONFIGS["water_ccpvdz_rhf"])
    r = ctx.post_hf(case["ao"], case["C"], None, case["eps"], None, case["no"], case["nv"], stages=2, maxiter=200, cc_conv=1e-12)
    assert r["iters"] == g["water_ccpvdz_rhf"]["cc_iters"]


def test_amplitude_restart_continues_the_same_trajectory(ctx):
    """Opt-in restart (qcp2_ccsd_set_start_T1 + T2 as the guess): stopping after 10 iterations and resuming must continue the very
    trajectory of the uninterrupted solve -- same energies from iteration 10 on, same total iteration count."""
    case = molecule_case(**CONFIGS["water_ccpvdz_rhf"])
    no, nv = case["no"], case["nv"]
    e_mp2, so, T2g = ctx.MP2(case["ao"], case["C"], None, case["eps"], case["eps"], no, nv)
    foo, fov, fvv = fock_blocks(case["eps"], case["eps"], no)
    full = ctx.CCSD(foo, fov, fvv, T2g, 200, 1e-12, so_ints=so)
    part = ctx.CCSD(foo, fov, fvv, T2g, 10, 1e-12, so_ints=so)
    assert part["iters"] == 10
    ctx.ccsd_set_start_T1(part["T1"])
    rest = ctx.CCSD(foo, fov, fvv, part["T2"], 200, 1e-12, so_ints=so)
    assert rest["iters"] == full["iters"] - 10
    assert np.max(np.abs(rest["e_hist"] - full["e_hist"][10:])) < 1e-13
    assert abs(rest["ccsd_energy"] - full["ccsd_energy"]) < 1e-13
    again = ctx.CCSD(foo, fov, fvv, T2g, 200, 1e-12, so_ints=so)      # the setting was one-shot
    assert again["iters"] == full["iters"]


def test_config3_ch4_ccpvtz_fused_pipeline_vs_committed_cpu_oracle(ctx):
    """BASELINE.json config 3 (o=10, v=162): the CPU side takes ~15 min, so its numbers are committed
    (tests/golden/ch4_ccpvtz_oracle.json, generator beside it).  The host SCF is deterministic, so
    the GPU box reproduces the same boundary inputs."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ch4_ccpvtz_oracle.json")))
    case = molecule_case("methane.xyz", "cc-pvtz", conv=1e-10, maxiter=400)
    assert (case["n"], case["no"], case["nv"]) == (86, 10, 162)
    assert abs(case["e_scf"] - g["e_scf"]) < 1e-9
    r = ctx.post_hf(case["ao"], case["C"], None, case["eps"], None, case["no"], case["nv"], stages=3, maxiter=200, cc_conv=1e-12)
    assert abs(r["e_mp2"] - g["e_mp2"]) < E_TOL and abs(r["e_ccsd"] - g["e_ccsd"]) < E_TOL and abs(r["e_t"] - g["e_t"]) < E_TOL
    assert r["iters"] == g["iters"] and np.max(np.abs(r["e_hist"] - np.array(g["e_hist"]))) < E_TOL
    assert ctx.ccsd_last_mode() == 2
    # ... and against the reference's OWN MP2() / CCSD() on this config (tests/golden/make_reference_golden_ch4.py: 37 minutes
    # of the reference's code on the shim's OpenMP GEMM; its CCSD_T() cannot hold 4 x o^3 v^3 = 136 GB)
    gr = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_energies.json")))["ch4_ccpvtz_rhf"]
    assert abs(r["e_mp2"] - gr["e_mp2"]) < E_TOL and abs(r["e_ccsd"] - gr["e_ccsd"]) < E_TOL and r["iters"] == gr["cc_iters"]
    # the same solve with the spin classes switched off (one class, dense pairs): identical trajectory to rounding
    os.environ["QCP2_CCSD_NO_SPIN_BLOCKS"] = "1"
    try:
        d = ctx.post_hf(case["ao"], case["C"], None, case["eps"], None, case["no"], case["nv"], stages=2, maxiter=200, cc_conv=1e-12)
    finally:
        del os.environ["QCP2_CCSD_NO_SPIN_BLOCKS"]
    assert ctx.ccsd_last_mode() == 1 and d["iters"] == r["iters"] and abs(d["e_ccsd"] - r["e_ccsd"]) < 1e-12


def test_p2_driver_runs_reference_input_json(tmp_path):
    """The C++ host mirror end to end: bin/P2 on the reference's own input.json (water/STO-3G CCSD)
    and on a CCSD(T) variant of it; energies vs the Crawford known answers."""
    exe = os.path.join(ROOT, "qc-p2_b200", "bin", "P2")
    assert os.path.exists(exe), "build qc-p2_b200/host first (__graft_entry__.build())"
    cfg = json.load(open(os.path.join(ROOT, "input.json")))
    cfg["type"] = "CCSD(T)"
    inp = tmp_path / "input.json"
    inp.write_text(json.dumps(cfg))
    out = subprocess.run([exe, str(inp)], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    r = json.loads(line[len("RESULT "):])
    assert abs(r["e_scf"] - (-74.942079928192)) < 1e-9
    assert abs(r["e_mp2"] - (-0.049149636120)) < 1e-10
    assert abs(r["e_ccsd"] - (-0.070680088376)) < 1e-10
    assert abs(r["e_t"] - (-0.000099877272)) < 1e-11
    assert "CC energy converged" in out.stdout
    fused = subprocess.run([exe, "--fused", str(inp)], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert fused.returncode == 0, fused.stderr[-2000:]
    rf = json.loads([l for l in fused.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    for key in ("e_mp2", "e_ccsd", "e_t"):
        assert abs(rf[key] - r[key]) < 1e-13, key
