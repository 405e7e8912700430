"""GPU parity tests (run with -m gpu on a B200): every C-ABI entry point against the CPU
oracles on the same seeded inputs.  Tolerances: energies 1e-10 Eh absolute (the bar of
BASELINE.json:north_star; observed agreement is ~1e-15), tensors 1e-12 relative to their
largest element.  All calls go through include/qcp2_b200.h via qc-p2_b200/qcp2.py."""
import json
import os

import numpy as np
import pytest
from helpers import fock_blocks, oracle_pipeline, rhf_case, uhf_case

import qcp2_b200
from oracle import np_oracle as npo
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
E_TOL = 1e-10
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "energies.json")))


def close(a, b, rel=1e-12):
    scale = max(1.0, float(np.max(np.abs(b)))) if np.size(b) else 1.0
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) <= rel * scale if np.size(b) else True


# ---------------------------------------------------------------------------- GEMM / contract
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (64, 64, 16), (65, 63, 17), (100, 26244 // 4, 33), (38, 1444, 144),
                                   (400, 1000, 88), (130, 257, 1320), (1, 300, 70), (300, 1, 70), (81, 39, 1521),
                                   (100, 2000, 5000), (10, 162, 40000), (162, 162, 16200), (3, 5, 70000),  # split-K / 112-row tile shapes
                                   (10, 5001, 162), (5001, 10, 162), (16, 700, 35), (701, 16, 33),   # skinny 16x256 / 256x16 tiles
                                   (45, 1001, 3001), (48, 129, 2000), (33, 257, 513),                 # 48x128 tile, odd strides
                                   (10, 10, 50001), (1, 3000, 1621), (3000, 1, 1621), (9, 39, 13689)])  # one-wave split-K factors
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_dgemm_matches_numpy(ctx, M, N, K, tA, tB):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((K, M) if tA else (M, K))
    B = rng.standard_normal((N, K) if tB else (K, N))
    C0 = rng.standard_normal((M, N))
    got = ctx.dgemm(tA, tB, 0.75, A, B, -0.5, C0)
    want = 0.75 * ((A.T if tA else A) @ (B.T if tB else B)) - 0.5 * C0
    assert np.max(np.abs(got - want)) < 1e-12 * max(1.0, K ** 0.5)


def test_contract_with_large_contracted_dimension_uses_split_k(ctx):
    """T1-residual pattern cc.cpp:276 at a size where the output is 10 x 60 and K = 10*60*60."""
    rng = np.random.default_rng(11)
    o, v = 10, 60
    td, ovvv = rng.standard_normal((o, o, v, v)), rng.standard_normal((o, v, v, v))
    got = ctx.contract(-0.5, td, "imef", ovvv, "maef", 0.0, np.zeros((o, v)), "ia")
    want = -0.5 * np.einsum("imef,maef->ia", td, ovvv)
    assert np.max(np.abs(got - want)) < 1e-10


def test_dgemm_large_tiles(ctx):
    rng = np.random.default_rng(1)
    for (M, N, K) in [(512, 4096, 300), (400, 8192, 1344), (1000, 1000, 1000)]:
        A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
        got = ctx.dgemm(0, 0, 1.0, A, B, 0.0, np.zeros((M, N)))
        assert np.max(np.abs(got - A @ B)) < 1e-11 * K ** 0.5
        got = ctx.dgemm(1, 1, 1.0, np.ascontiguousarray(A.T), np.ascontiguousarray(B.T), 0.0, np.zeros((M, N)))
        assert np.max(np.abs(got - A @ B)) < 1e-11 * K ** 0.5


CONTRACT_PATTERNS = [
    ("pqrs", "sl", "pqrl"), ("pqrl", "rk", "pqkl"), ("qj", "pqkl", "pjkl"), ("pi", "pjkl", "ijkl"),
    ("me", "ma", "ae"), ("mf", "mafe", "ae"), ("mnaf", "mnef", "ae"), ("ie", "me", "mi"), ("ne", "mnie", "mi"),
    ("inef", "mnef", "mi"), ("nf", "mnef", "me"), ("je", "mnie", "mnij"), ("ie", "mnje", "mnij"),
    ("ijef", "mnef", "mnij"), ("mb", "amef", "abef"), ("ma", "bmef", "abef"), ("mnab", "mnef", "abef"),
    ("jf", "mbef", "mbej"), ("nb", "mnej", "mbej"), ("jnfb", "mnef", "mbej"), ("jfnb", "mnef", "mbej"),
    ("ie", "ae", "ia"), ("ma", "mi", "ia"), ("imae", "me", "ia"), ("nf", "naif", "ia"), ("imef", "maef", "ia"),
    ("mnae", "nmei", "ia"), ("ijae", "be", "ijab"), ("ijbe", "ae", "ijab"), ("mb", "me", "be"),
    ("imab", "mj", "ijab"), ("jmab", "mi", "ijab"), ("je", "me", "jm"), ("imab", "jm", "ijab"),
    ("mnab", "mnij", "ijab"), ("ijef", "abef", "ijab"), ("imae", "mbej", "ijab"), ("jmae", "mbei", "ijab"),
    ("imbe", "maej", "ijab"), ("jmbe", "maei", "ijab"), ("iema", "mbje", "ijab"), ("jemb", "maie", "ijab"),
    ("ie", "abej", "ijab"), ("je", "abei", "ijab"), ("ma", "mbij", "ijab"), ("mb", "maij", "ijab"),
    ("jkae", "eibc", "ijkabc"), ("jkce", "eiba", "ijkabc"), ("imbc", "majk", "ijkabc"), ("kmba", "mcji", "ijkabc")]


@pytest.mark.parametrize("pattern", CONTRACT_PATTERNS)
def test_contract_every_reference_pattern(ctx, pattern):
    """Every distinct index pattern of the reference's 71 btas::contract call sites."""
    ia, ib, ic = pattern
    rng = np.random.default_rng(sum(ord(ch) * (q + 1) for q, ch in enumerate(''.join(pattern))))
    occ, vir = 5, 9
    ext = {l: (occ if l in "ijkmn" else vir) for l in "abcefijklmnpqrs"}
    ext.update({l: 6 for l in "pqrsl"})
    ext["k"] = 6 if "p" in ia + ib + ic else occ
    ext["j"] = 6 if "p" in ia + ib + ic else occ
    ext["i"] = 6 if "p" in ia + ib + ic else occ
    A = rng.standard_normal([ext[l] for l in ia])
    B = rng.standard_normal([ext[l] for l in ib])
    C0 = rng.standard_normal([ext[l] for l in ic])
    got = ctx.contract(0.7, A, ia, B, ib, -0.3, C0, ic)
    want = -0.3 * C0 + 0.7 * np.einsum(f"{ia},{ib}->{ic}", A, B)
    assert np.max(np.abs(got - want)) < 1e-12
    assert np.max(np.abs(got - orc.contract(0.7, A, ia, B, ib, -0.3, C0, ic))) < 1e-12


# ---------------------------------------------------------------------------- integrals / mp2
@pytest.mark.parametrize("n", [1, 2, 7, 24, 33])
def test_transform_ao_mo(ctx, n):
    rng = np.random.default_rng(n)
    ao, C1, C2 = rng.standard_normal((n,) * 4), rng.standard_normal((n, n)), rng.standard_normal((n, n))
    got = ctx.transform_ao_mo(ao, C1, C2)
    assert close(got, orc.transform_ao_mo(ao, C1, C2))
    assert close(got, npo.transform_ao_mo(ao, C1, C2), 1e-11)


@pytest.mark.parametrize("n", [1, 3, 8])
def test_transform_to_so_slice_denom(ctx, n):
    rng = np.random.default_rng(100 + n)
    aa, bb, ab = (rng.standard_normal((n,) * 4) for _ in range(3))
    so = ctx.transform_to_so(aa, bb, ab)
    assert np.array_equal(so, orc.transform_to_so(aa, bb, ab))       # pure gather: bit-exact
    so_r = ctx.transform_to_so(aa, aa, aa)
    assert np.array_equal(so_r, orc.transform_to_so(aa, aa, aa))
    no = 2 * (n // 2) if n > 1 else 2
    nv = 2 * n - no
    for spec in qcp2_b200.INT_NAMES + ["vvoo", "vooo"]:
        assert np.array_equal(ctx.slice_ints(so, no, nv, spec), orc.slice_ints(so, no, nv, spec))
    ea, eb = np.sort(rng.standard_normal(n)), np.sort(rng.standard_normal(n))
    F = ctx.make_so_moes(ea, eb)
    assert np.array_equal(F, orc.make_so_moes(ea, eb))
    assert np.array_equal(ctx.make_denom(F, no, nv), orc.make_denom(F, no, nv))
    f1 = ctx.make_fock(ea, eb, 0, no) if no > 0 else None
    if no > 0 and no % 2 == 0:
        assert np.array_equal(f1, orc.make_fock(ea, eb, 0, no))
        assert np.array_equal(ctx.make_fock(ea, eb, no, 2 * n), orc.make_fock(ea, eb, no, 2 * n)) if nv > 0 else True


def test_slice_rejects_bad_spec(ctx):
    with pytest.raises(qcp2_b200.QCP2Error):
        ctx.slice_ints(np.zeros((4,) * 4), 2, 2, "oxvv")


@pytest.mark.parametrize("case_fn", [lambda: rhf_case(6, 2, 31), lambda: uhf_case(5, 3, 2, 32)])
def test_mp2_driver(ctx, case_fn):
    case = case_fn()
    r = oracle_pipeline(case, with_t=False, maxiter=1)
    uhf = "Ca" in case
    e, so, T = ctx.MP2(case["ao"], case["Ca"] if uhf else case["C"], case.get("Cb"), r["eps_a"], r["eps_b"],
                       case["no"], case["nv"], uhf=uhf)
    assert abs(e - r["e_mp2"]) < E_TOL
    assert close(so, r["so"]) and close(T, r["T2_mp2"])
    T_b, e_b = ctx.run_mp2(r["ints"]["oovv"], orc.make_denom(r["F"], case["no"], case["nv"]))
    assert abs(e_b - r["e_mp2"]) < E_TOL and close(T_b, r["T2_mp2"])


# ---------------------------------------------------------------------------- CCSD pieces
@pytest.fixture(scope="module")
def cc_state():
    """Mid-iteration CCSD state (non-zero T1) of an RHF and an as-written UHF pseudo-molecule."""
    out = {}
    for name, case in (("rhf", rhf_case(6, 2, 41)), ("uhf", uhf_case(5, 3, 2, 42))):
        r = oracle_pipeline(case, with_t=False, maxiter=3)
        out[name] = (case, r)
    return out


@pytest.mark.parametrize("kind", ["rhf", "uhf"])
def test_elementwise_pieces(ctx, cc_state, kind):
    case, r = cc_state[kind]
    assert close(ctx.make_tau(r["T1"], r["T2"]), orc.make_tau(r["T1"], r["T2"]), 1e-15)
    assert close(ctx.make_tau_bar(r["T1"], r["T2"]), orc.make_tau(r["T1"], r["T2"], True), 1e-15)
    assert close(ctx.multiply_Ts(r["T1"]), orc.multiply_Ts(r["T1"]), 1e-15)
    eo, ev = np.diag(r["foo"]), np.diag(r["fvv"])
    assert np.array_equal(ctx.make_D_ia(r["foo"], r["fvv"]), eo[:, None] - ev[None, :])
    D2 = ctx.make_D_ijab(r["foo"], r["fvv"])
    assert np.array_equal(D2, ((eo[:, None, None, None] + eo[None, :, None, None]) - ev[None, None, :, None]) - ev[None, None, None, :])
    D3 = ctx.make_D_triples(r["foo"], r["fvv"])
    assert D3.shape == (case["no"],) * 3 + (case["nv"],) * 3
    i, j, k, a, b, c = 1, 0, 2, 3, 1, 0
    assert D3[i, j, k, a, b, c] == eo[i] + eo[j] + eo[k] - ev[a] - ev[b] - ev[c]


@pytest.mark.parametrize("kind", ["rhf", "uhf"])
def test_intermediates_T1_T2_energy(ctx, cc_state, kind):
    case, r = cc_state[kind]
    I, foo, fov, fvv = r["ints"], r["foo"], r["fov"], r["fvv"]
    X_ref = orc.update_intermediates(r["T1"], r["T2"], I, foo, fov, fvv)
    X = ctx.update_intermediates(r["T1"], r["T2"], I, foo, fov, fvv)
    for k in qcp2_b200.INTER_NAMES:
        assert close(X[k], X_ref[k]), k
    X_none = ctx.update_intermediates(r["T1"], r["T2"], I, foo, None, fvv)      # f_ov = NULL means zeros (Q4)
    for k in qcp2_b200.INTER_NAMES:
        assert close(X_none[k], X_ref[k]), k
    fov_r = np.random.default_rng(5).standard_normal(fov.shape) * 0.01           # non-zero F_ia exercises cc.cpp:194,207,223,280,396
    Xr_ref = orc.update_intermediates(r["T1"], r["T2"], I, foo, fov_r, fvv)
    Xr = ctx.update_intermediates(r["T1"], r["T2"], I, foo, fov_r, fvv)
    for k in qcp2_b200.INTER_NAMES:
        assert close(Xr[k], Xr_ref[k]), k
    t1_ref = orc.make_T1(r["T1"], r["T2"], I, Xr_ref, foo, fov_r, fvv)
    assert close(ctx.make_T1(r["T1"], r["T2"], I, Xr_ref, foo, fov_r, fvv), t1_ref)
    t2_ref = orc.make_T2(t1_ref, r["T2"], I, X_ref, foo, fvv)
    assert close(ctx.make_T2(t1_ref, r["T2"], I, X_ref, foo, fvv), t2_ref)
    e_ref = orc.ccsd_energy(r["T1"], r["T2"], I["oovv"], fov_r)
    assert abs(ctx.ccsd_energy(r["T1"], r["T2"], I["oovv"], fov_r) - e_ref) < 1e-13
    if kind == "rhf":   # independent einsum formulation of the same equations
        In = npo.get_integrals(r["so"], case["no"])
        Xn = npo.intermediates(r["T1"], r["T2"], In, foo, fov, fvv)
        for k in qcp2_b200.INTER_NAMES:
            assert close(X[k], Xn[k], 1e-11), k


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_full_pipeline_against_golden(ctx, name):
    g = GOLDEN[name]
    spec = g["spec"]
    case = rhf_case(spec["n"], spec["ndocc"], spec["seed"]) if spec["kind"] == "rhf" else \
        uhf_case(spec["n"], spec["na"], spec["nb"], spec["seed"])
    uhf = spec["kind"] == "uhf"
    no, nv = case["no"], case["nv"]
    ea, eb = (case["eps_a"], case["eps_b"]) if uhf else (case["eps"], case["eps"])
    e_mp2, so, T2g = ctx.MP2(case["ao"], case["Ca"] if uhf else case["C"], case.get("Cb"), ea, eb, no, nv, uhf=uhf)
    assert abs(e_mp2 - g["e_mp2"]) < E_TOL
    foo, fov, fvv = fock_blocks(ea, eb, no)
    res = ctx.CCSD(foo, fov, fvv, T2g, 100, 1e-12, so_ints=so, want_sliced=True)
    assert res["iters"] == g["iters"]
    assert np.max(np.abs(res["e_hist"] - np.array(g["e_hist"]))) < E_TOL
    assert abs(res["ccsd_energy"] - g["e_ccsd"]) < E_TOL
    S = res["sliced_ints"]
    eo, ev = np.diag(foo).copy(), np.diag(fvv).copy()
    mode = ctx.t_select_mode(res["T2"], S["oovv"], S["vovv"], S["ovoo"])
    assert mode == (qcp2_b200.T_LITERAL if uhf else qcp2_b200.T_SYMMETRIC)
    e_t = ctx.CCSD_T(res["T1"], res["T2"], S["oovv"], S["vovv"], S["ovoo"], eo, ev)
    assert abs(e_t - g["e_t"]) < E_TOL
    e_lit = ctx.CCSD_T(res["T1"], res["T2"], S["oovv"], S["vovv"], S["ovoo"], eo, ev, qcp2_b200.T_LITERAL)
    assert abs(e_lit - g["e_t"]) < E_TOL


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_fused_post_hf_pipeline_against_golden(ctx, name):
    g = GOLDEN[name]
    spec = g["spec"]
    uhf = spec["kind"] == "uhf"
    case = uhf_case(spec["n"], spec["na"], spec["nb"], spec["seed"]) if uhf else rhf_case(spec["n"], spec["ndocc"], spec["seed"])
    ea, eb = (case["eps_a"], case["eps_b"]) if uhf else (case["eps"], None)
    r = ctx.post_hf(case["ao"], case["Ca"] if uhf else case["C"], case.get("Cb"), ea, eb, case["no"], case["nv"], uhf=uhf,
                    stages=3, maxiter=100, cc_conv=1e-12)
    assert abs(r["e_mp2"] - g["e_mp2"]) < E_TOL and abs(r["e_ccsd"] - g["e_ccsd"]) < E_TOL and abs(r["e_t"] - g["e_t"]) < E_TOL
    assert r["iters"] == g["iters"] and np.max(np.abs(r["e_hist"] - np.array(g["e_hist"]))) < E_TOL
    r1 = ctx.post_hf(case["ao"], case["Ca"] if uhf else case["C"], case.get("Cb"), ea, eb, case["no"], case["nv"], uhf=uhf, stages=1)
    assert abs(r1["e_mp2"] - g["e_mp2"]) < E_TOL and r1["e_ccsd"] == 0.0


def test_ccsd_from_presliced_blocks_and_maxiter_cap(ctx):
    case = rhf_case(6, 2, 51)
    r = oracle_pipeline(case, with_t=False, maxiter=4)          # stops at the cap, not converged
    res = ctx.CCSD(r["foo"], r["fov"], r["fvv"], r["T2_mp2"], 4, 1e-12, integrals=r["ints"])
    assert res["iters"] == r["iters"] == 4
    assert np.max(np.abs(res["e_hist"] - r["e_hist"])) < E_TOL
    assert close(res["T2"], r["T2"]) and close(res["T1"], r["T1"])


def test_so_block_is_the_slice_of_transform_to_so_bit_for_bit(ctx):
    """The fused pipeline gathers every o/v block straight from the spatial MO tensors (the (2n)^4 tensor is never built);
    the values must be the ones integrals.cpp:136-167 + 184-207 produce: RHF-type (aa = bb = ab) and as-written UHF."""
    rng = np.random.default_rng(5)
    n, o = 5, 4
    v = 2 * n - o
    aa, bb, ab = (rng.standard_normal((n,) * 4) for _ in range(3))
    for (x, y, z) in ((aa, aa, aa), (aa, bb, ab)):
        so = npo.transform_to_so(x, y, z)
        for spec in qcp2_b200.INT_NAMES:
            got = ctx.so_block(x, y, z, o, v, spec)
            assert np.array_equal(got, npo.slice_ints(so, o, spec)), spec


# ---------------------------------------------------------------------------- (T)
@pytest.mark.parametrize("o,v", [(3, 4), (4, 6), (4, 7), (5, 17), (6, 33), (3, 39)])  # v(v-1)/2 odd for v = 6, 7; v^2 odd for 7, 17, 33, 39
def test_t_synthetic_small_all_modes(ctx, o, v):
    s = npo.synthetic_t_inputs(o, v)
    args = (s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"])
    ref = orc.ccsd_t(s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], np.diag(s["eps_o"]), np.diag(s["eps_v"]))
    for mode in (qcp2_b200.T_LITERAL, qcp2_b200.T_SYMMETRIC, qcp2_b200.T_AUTO):
        got = ctx.CCSD_T(*args, mode)
        assert abs(got - ref) < E_TOL and abs(got - ref) < 1e-12 * abs(ref), (mode, got, ref)
    assert ctx.t_select_mode(s["T2"], s["oovv"], s["vovv"], s["ovoo"]) == qcp2_b200.T_SYMMETRIC


def test_t_tma_kernel_serves_every_v_and_both_modes(ctx):
    """v = 162 (CH4/cc-pVTZ: v(v-1)/2 = 13 041 is odd) and the literal mode (N = v^2) run on the TMA/mbarrier kernel: the
    packed rows are padded to an even stride.  Checked against the per-triple numpy restatement."""
    import torch
    o, v = 4, 162
    s = npo.synthetic_t_inputs(o, v)
    names = ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")
    ref = npo.ccsd_t_blocked(*(s[k] for k in names), True)
    got = ctx.CCSD_T(*(s[k] for k in names), qcp2_b200.T_SYMMETRIC)
    assert abs(got - ref) < 1e-12 * abs(ref), (got, ref)
    got_lit = ctx.CCSD_T(*(s[k] for k in names), qcp2_b200.T_LITERAL)
    assert abs(got_lit - ref) < 1e-11 * abs(ref), (got_lit, ref)
    ctx.set_pointer_mode(qcp2_b200.DEVICE)
    try:
        d = {k: torch.tensor(s[k], device="cuda") for k in names}
        for mode, ncols in ((qcp2_b200.T_SYMMETRIC, v * (v - 1) // 2), (qcp2_b200.T_LITERAL, v * v)):
            plan = qcp2_b200.DeviceTriples(ctx, *(d[k] for k in names), mode)
            info = plan.info()
            assert info["uses_tma_kernel"] and info["ncols"] == ncols and info["row_stride"] % 2 == 0
            e = plan.run(0, 1, 3)
            assert np.isfinite(e)
            plan.close()
    finally:
        ctx.set_pointer_mode(qcp2_b200.HOST)


def test_t_no_triples_and_unsymmetric_inputs(ctx):
    s = npo.synthetic_t_inputs(2, 3)
    args = (s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"])
    assert ctx.CCSD_T(*args, qcp2_b200.T_SYMMETRIC) == 0.0           # no i<j<k exists
    rng = np.random.default_rng(3)
    o, v = 4, 5
    T1, T2 = rng.standard_normal((o, v)) * 0.1, rng.standard_normal((o, o, v, v)) * 0.1      # no symmetry at all
    oovv, vovv, ovoo = (rng.standard_normal(sh) * 0.1 for sh in ((o, o, v, v), (v, o, v, v), (o, v, o, o)))
    eo, ev = -1.0 - rng.random(o), 1.0 + rng.random(v)
    assert ctx.t_select_mode(T2, oovv, vovv, ovoo) == qcp2_b200.T_LITERAL
    ref = orc.ccsd_t(T1, T2, oovv, vovv, ovoo, np.diag(eo), np.diag(ev))
    assert abs(ctx.CCSD_T(T1, T2, oovv, vovv, ovoo, eo, ev) - ref) < 1e-12 * max(1.0, abs(ref))


def test_t_reduced_synthetic_shape_full_sum(ctx):
    """SURVEY.md section 8d reduced shape o=12, v=96: full E(T) vs the per-triple numpy oracle."""
    o, v = 12, 96
    s = npo.synthetic_t_inputs(o, v)
    args = (s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"])
    bu, cu = np.triu_indices(v, 1)
    trip = npo.sorted_triples(o)
    per = np.array([npo.ccsd_t_triple_packed_np(*args, int(i), int(j), int(k), bu, cu) for i, j, k in trip])
    e_tr = np.zeros(len(trip))
    got = ctx.CCSD_T_range(*args, qcp2_b200.T_SYMMETRIC, 0, 1, -1, e_tr)
    assert np.max(np.abs(e_tr - per)) < 1e-12 * np.max(np.abs(per))
    assert abs(got - per.sum()) < E_TOL
    # sharding property: strided ranges partition the sum exactly (per-triple values are identical)
    e2 = np.zeros(len(trip))
    parts = [ctx.CCSD_T_range(*args, qcp2_b200.T_SYMMETRIC, r, 3, -1, e2) for r in range(3)]
    assert np.array_equal(e2, e_tr)
    assert abs(sum(parts) - got) < 1e-12 * abs(got)


def test_t_full_size_sampled_triples_and_device_plan():
    """Config 5 (o=40, v=400) at full size: sampled triples vs the numpy oracle on identical
    (device-generated, host-verified) inputs; size-independent properties on the rest."""
    import torch
    ctx = qcp2_b200.Context(0)
    o, v = 40, 400
    d = qcp2_b200.synth_t_inputs_device(ctx, o, v)
    host = {k: t.cpu().numpy() for k, t in d.items()}
    # generator parity: spot-check the device tensors against the numpy generator on slices
    small = npo.synthetic_t_inputs(6, 9)
    dsmall = qcp2_b200.synth_t_inputs_device(ctx, 6, 9)
    for k in small:
        assert np.array_equal(dsmall[k].cpu().numpy(), small[k]), k
    plan = qcp2_b200.DeviceTriples(ctx, d["T1"], d["T2"], d["oovv"], d["vovv"], d["ovoo"], d["eps_o"], d["eps_v"],
                                   qcp2_b200.T_SYMMETRIC)
    nt = ctx.t_count(o, qcp2_b200.T_SYMMETRIC)
    assert nt == 9880
    e_dev = torch.zeros(nt, dtype=torch.float64, device="cuda")
    trip = npo.sorted_triples(o)
    bu, cu = np.triu_indices(v, 1)
    args = tuple(host[k] for k in ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v"))
    samples = [0, 37, 9879, 4321]          # (0,1,2), (0,1,39), (37,38,39), interior
    assert tuple(trip[37]) == (0, 1, 39) and tuple(trip[9879]) == (37, 38, 39)
    for t in samples:
        plan.run(t, 1, 1, e_dev)
    got = e_dev.cpu().numpy()
    for t in samples:
        i, j, k = map(int, trip[t])
        ref = npo.ccsd_t_triple_packed_np(*args, i, j, k, bu, cu)
        assert abs(got[t] - ref) <= 1e-12 * abs(ref), (t, got[t], ref)
    # a strided batch reproduces the single-triple values bit for bit (batching independence)
    e2 = torch.zeros(nt, dtype=torch.float64, device="cuda")
    part = plan.run(0, 4321, 3, e2)   # triples 0, 4321, 8642
    g2 = e2.cpu().numpy()
    assert g2[0] == got[0] and g2[4321] == got[4321]
    assert abs(part - (g2[0] + g2[4321] + g2[8642])) <= 1e-15 * abs(part)
    tm = plan.timings()
    assert tm["gemm_launches"] >= 1 and tm["gemm_seconds"] > 0
    plan.close()
    ctx.close()


def test_t_streamed_and_resident_vovv_paths_agree(ctx, monkeypatch):
    """Host-pointer calls stream <vo||vv> slab by slab (H2D + packing overlapped with the triples, colex
    schedule); QCP2_T_NO_STREAM=1 takes the resident path.  Same per-triple values and sums, bit for bit;
    AUTO falls back to LITERAL when only <vo||vv> breaks the antisymmetry (caught by the slab-wise check)."""
    o, v = 7, 30
    s = npo.synthetic_t_inputs(o, v)
    args = [s[k] for k in ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")]
    nt = ctx.t_count(o, qcp2_b200.T_SYMMETRIC)
    e_stream, e_res = np.zeros(nt), np.zeros(nt)
    a = ctx.CCSD_T_range(*args, qcp2_b200.T_SYMMETRIC, 0, 1, -1, e_stream)
    b = ctx.CCSD_T_range(*args, qcp2_b200.T_SYMMETRIC, 1, 3, 5, None)
    monkeypatch.setenv("QCP2_T_NO_STREAM", "1")
    a2 = ctx.CCSD_T_range(*args, qcp2_b200.T_SYMMETRIC, 0, 1, -1, e_res)
    b2 = ctx.CCSD_T_range(*args, qcp2_b200.T_SYMMETRIC, 1, 3, 5, None)
    monkeypatch.delenv("QCP2_T_NO_STREAM")
    assert np.array_equal(e_stream, e_res) and a == a2 and b == b2
    bad = [x.copy() for x in args]
    bad[3][1, 2, 3, 4] += 0.25  # <vo||vv> no longer antisymmetric in (b,c); everything else still is
    ref = orc.ccsd_t(bad[0], bad[1], bad[2], bad[3], bad[4], np.diag(bad[5]), np.diag(bad[6]))
    got = ctx.CCSD_T(*bad, qcp2_b200.T_AUTO)
    assert abs(got - ref) < 1e-12 * max(1.0, abs(ref))
    assert abs(got - a) > 1e-8  # and it is a different number from the symmetric evaluation


def test_t_mgpu_entry_point_on_a_single_rank_communicator():
    """qcp2_ccsd_t_mgpu with the library's own NCCL communicator at world size 1 (what one GPU can
    check; tools/mgpu_check.py does ranks 2..8 under torchrun): same energy and per-triple values as
    qcp2_ccsd_t, bit for bit, in host and device pointer mode."""
    import torch
    c = qcp2_b200.Context(0)
    c.comm_init(qcp2_b200.comm_unique_id(), 1, 0)
    assert c.comm_rank() == (0, 1)
    o, v = 6, 33
    s = npo.synthetic_t_inputs(o, v)
    names = ("T1", "T2", "oovv", "vovv", "ovoo", "eps_o", "eps_v")
    args = tuple(s[k] for k in names)
    e_tr = np.zeros(c.t_count(o, qcp2_b200.T_SYMMETRIC))
    want = c.CCSD_T_range(*args, qcp2_b200.T_SYMMETRIC, 0, 1, -1, e_tr)
    got, per = c.CCSD_T_mgpu(*args, o, v, qcp2_b200.T_AUTO, root=0, want_triples=True)
    seq = 0.0
    for x in e_tr.tolist():  # (the builtin sum() compensates since Python 3.12)
        seq += x
    assert got == seq == want and np.array_equal(per[:len(e_tr)], e_tr)
    # device pointer mode + the collective building blocks
    c.set_pointer_mode(qcp2_b200.DEVICE)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in args]
    assert c.CCSD_T_mgpu(*dev, o, v, qcp2_b200.T_SYMMETRIC, root=0) == got
    x = torch.arange(5, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    c.comm_broadcast(x, 0), c.comm_allreduce_sum(x), c.comm_allgather(x, y)
    c.synchronize()
    assert torch.equal(x.cpu(), torch.arange(5, dtype=torch.float64)) and torch.equal(x, y)
    c.comm_destroy()
    assert c.comm_rank() == (0, 1)
    c.close()


def test_post_hf_mgpu_on_a_single_rank_communicator():
    """qcp2_post_hf_mgpu (collective fused pipeline) at world size 1: the same numbers as qcp2_post_hf; ranks 2..8 with
    sharded CCSD products are checked by tools/mgpu_ccsd_check.py under torchrun (profiles/r1_mgpu_ccsd_check_*)."""
    c = qcp2_b200.Context(0)
    c.comm_init(qcp2_b200.comm_unique_id(), 1, 0)
    for case in (rhf_case(6, 2, 81), uhf_case(5, 3, 2, 82)):
        uhf = "Ca" in case
        Ca = case["Ca"] if uhf else case["C"]
        ea, eb = (case["eps_a"], case["eps_b"]) if uhf else (case["eps"], None)
        a = c.post_hf(case["ao"], Ca, case.get("Cb"), ea, eb, case["no"], case["nv"], uhf=uhf, stages=3, maxiter=60, cc_conv=1e-12)
        b = c.post_hf_mgpu(case["ao"], Ca, case.get("Cb"), ea, eb, case["n"], case["no"], case["nv"], uhf=uhf, stages=3, maxiter=60,
                           cc_conv=1e-12, root=0)
        for k in ("e_mp2", "e_ccsd", "e_t"):
            assert a[k] == b[k], k
        assert a["iters"] == b["iters"] and np.array_equal(a["e_hist"], b["e_hist"])
    assert c.sharded_gemms == 0
    c.close()
