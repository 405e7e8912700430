"""Pins the CPU oracle against THE REFERENCE ITSELF (CPU, no GPU).

oracle/_ref/libqcp2_ref.so is the reference's own src/{integrals,mp2,cc,hf,general}.cpp, unmodified,
compiled by oracle/ref_shim/Makefile against API stand-ins for the absent third-party headers
(BTAS, Eigen, Libint2, nlohmann-json).  Every function of SURVEY.md section 8a is called on seeded
inputs through it and through the restatement in oracle/oracle.cpp; they have to agree to rounding.
The committed fixture tests/golden/reference_energies.json (generator beside it) carries the
reference's results for BASELINE.json configs 1, 2 and 4 to the GPU box, where /root/reference and
possibly oracle/_ref do not exist.
"""
import json
import os

import numpy as np
import pytest
from helpers import ROOT, fock_blocks, molecule_case, oracle_pipeline, rhf_case, uhf_case

from oracle import oracle as orc
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")

CASES = {
    "rhf_n5": lambda: rhf_case(5, 2, 11),
    "rhf_n6_odd_v": lambda: rhf_case(6, 1, 5),
    "uhf_n5_3a2b": lambda: uhf_case(5, 3, 2, 12),
    "uhf_n4_2a1b": lambda: uhf_case(4, 2, 1, 3),
}
TOL = 1e-13


def close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    assert float(np.max(np.abs(a - b))) <= tol * scale if a.size else True


@pytest.fixture(scope="module", params=sorted(CASES))
def staged(request):
    """Boundary inputs + every intermediate stage, computed by the ORACLE, so that each reference
    function can be fed exactly the operands the oracle's function saw."""
    case = CASES[request.param]()
    return case, oracle_pipeline(case, maxiter=60, conv=1e-12)


def amplitudes(o, v, seed):
    """Unsymmetric random T1/T2: nothing in the comparison may lean on permutational symmetry."""
    rng = np.random.default_rng(seed)
    return 0.1 * rng.standard_normal((o, v)), 0.05 * rng.standard_normal((o, o, v, v))


# every distinct index pattern of the reference's 71 btas::contract call sites (integrals.cpp:121-129, cc.cpp:194-523)
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
def test_shim_contract_is_einsum(pattern):
    """What stays restated in the compiled reference is BTAS' contract itself (the library is absent).  The shim's
    implementation (permute -> GEMM -> permute, like BTAS) against numpy.einsum and against the oracle's engine on
    every index pattern the reference uses, with alpha/beta accumulation."""
    ia, ib, ic = pattern
    rng = np.random.default_rng(sum(ord(ch) * (q + 1) for q, ch in enumerate("".join(pattern))))
    ext = {l: (4 if l in "ijkmn" else 6) for l in "abcefijklmnpqrs"}
    if "p" in ia + ib + ic:
        ext.update({l: 5 for l in "pqrslijk"})
    A, B, C0 = (rng.standard_normal([ext[l] for l in lab]) for lab in (ia, ib, ic))
    got = ref.shim_contract(0.7, A, ia, B, ib, -0.3, C0, ic)
    want = -0.3 * C0 + 0.7 * np.einsum(f"{ia},{ib}->{ic}", A, B)
    assert np.max(np.abs(got - want)) < 1e-13
    assert np.max(np.abs(got - orc.contract(0.7, A, ia, B, ib, -0.3, C0, ic))) < 1e-13


def test_transform_ao_mo_and_to_so(staged):
    case, st = staged
    uhf = "Ca" in case
    C1, C2 = (case["Ca"], case["Cb"]) if uhf else (case["C"], case["C"][:, ::-1].copy())
    close(ref.transform_ao_mo(case["ao"], C1, C2), orc.transform_ao_mo(case["ao"], C1, C2))  # asymmetric C roles
    close(ref.transform_ao_mo(case["ao"], C1, C1), st["mo_aa"])
    so = ref.transform_to_so(st["mo_aa"], st["mo_bb"], st["mo_ab"])
    assert np.array_equal(so, st["so"])  # pure gather / subtraction of the same operands: bit-exact


def test_slices_moes_denominators_mp2(staged):
    case, st = staged
    no, nv = case["no"], case["nv"]
    for spec in ref.INT_NAMES:
        assert np.array_equal(ref.slice_ints(st["so"], no, nv, spec), st["ints"][spec]), spec
    F = ref.make_so_moes(st["eps_a"], st["eps_b"])
    assert np.array_equal(F, st["F"])
    den = ref.make_denom(F, no, nv)
    assert np.array_equal(den, orc.make_denom(F, no, nv))
    T, e = ref.run_mp2(st["ints"]["oovv"], den)
    To, eo = orc.run_mp2(st["ints"]["oovv"], den)
    assert np.array_equal(T, To) and abs(e - eo) <= 1e-15 * max(1.0, abs(eo))
    assert abs(e - st["e_mp2"]) < 1e-14


def test_fock_blocks_and_denominators(staged):
    case, st = staged
    no, nv = case["no"], case["nv"]
    n2 = no + nv
    ea, eb = st["eps_a"], st["eps_b"]
    for n1, n2_ in ((0, n2), (0, no)) + (((no, n2),) if no % 2 == 0 else ()):  # odd n1 branch is wrong AND unreachable (Q6)
        assert np.array_equal(ref.make_fock(ea, eb, n1, n2_), orc.make_fock(ea, eb, n1, n2_))
    foo, _, fvv = fock_blocks(ea, eb, no)
    assert np.array_equal(ref.make_fock(ea, eb, 0, no), foo)
    eo, ev = np.diag(foo), np.diag(fvv)
    assert np.array_equal(ref.make_D(foo, fvv, 0), eo[:, None] - ev[None, :])
    D2 = ref.make_D(foo, fvv, 1)
    assert np.array_equal(D2, (eo[:, None, None, None] + eo[None, :, None, None]) - ev[None, None, :, None] - ev[None, None, None, :])
    D3 = ref.make_D(foo, fvv, 2)
    want = (eo[:, None, None, None, None, None] + eo[None, :, None, None, None, None] + eo[None, None, :, None, None, None]
            - ev[None, None, None, :, None, None] - ev[None, None, None, None, :, None] - ev[None, None, None, None, None, :])
    close(D3, want, 1e-15)


def test_tau_family(staged):
    case, _ = staged
    Ts, Td = amplitudes(case["no"], case["nv"], 21)
    assert np.array_equal(ref.make_tau(Ts, Td), orc.make_tau(Ts, Td))
    assert np.array_equal(ref.make_tau(Ts, Td, bar=True), orc.make_tau(Ts, Td, bar=True))
    assert np.array_equal(ref.multiply_Ts(Ts), orc.multiply_Ts(Ts))


def test_intermediates_and_residuals(staged):
    """update_intermediates (17 contract calls), make_T1 (6), make_T2 (26) of the reference vs the
    restatement, on unsymmetric amplitudes and a non-zero F_ia (so no term can hide)."""
    case, st = staged
    no, nv = case["no"], case["nv"]
    Ts, Td = amplitudes(no, nv, 33)
    fov = 0.01 * np.random.default_rng(5).standard_normal((no, nv))
    I, foo, fvv = st["ints"], st["foo"], st["fvv"]
    Xr = ref.update_intermediates(Ts, Td, I, foo, fov, fvv)
    Xo = orc.update_intermediates(Ts, Td, I, foo, fov, fvv)
    for k in ref.INTER_NAMES:
        close(Xr[k], Xo[k])
    close(ref.make_T1(Ts, Td, I, Xo, foo, fov, fvv), orc.make_T1(Ts, Td, I, Xo, foo, fov, fvv))
    close(ref.make_T2(Ts, Td, I, Xo, foo, fvv), orc.make_T2(Ts, Td, I, Xo, foo, fvv))
    er, eo = ref.ccsd_energy(Ts, Td, I["oovv"], fov), orc.ccsd_energy(Ts, Td, I["oovv"], fov)
    assert abs(er - eo) <= 1e-14 * max(1.0, abs(eo))


def test_ccsd_t_literal(staged):
    """CCSD_T (cc.cpp:454-560: 9 disconnected terms, 18 contract calls) vs the restatement AND vs the
    per-triple X = A.B form the GPU kernel implements, on unsymmetric amplitudes."""
    from oracle import np_oracle as npo
    case, st = staged
    no, nv = case["no"], case["nv"]
    Ts, Td = amplitudes(no, nv, 44)
    I, foo, fvv = st["ints"], st["foo"], st["fvv"]
    e_ref = ref.ccsd_t(Ts, Td, I["oovv"], I["vovv"], I["ovoo"], foo, fvv)
    e_orc = orc.ccsd_t(Ts, Td, I["oovv"], I["vovv"], I["ovoo"], foo, fvv)
    assert abs(e_ref - e_orc) <= 1e-13 * max(1.0, abs(e_orc))
    e_blk = orc.ccsd_t_blocked(Ts, Td, I["oovv"], I["vovv"], I["ovoo"], np.diag(foo).copy(), np.diag(fvv).copy(), npo.all_triples(no))
    assert abs(e_ref - e_blk) <= 1e-12 * max(1.0, abs(e_ref))


def test_whole_chain_iteration_count_and_energies(staged):
    """MP2 body -> CCSD() -> make_moe_tensors -> CCSD_T() of the reference (UHF: as written, with the
    empty F_ia read as zeros, Q4/Q5) vs the oracle pipeline: same energies, same number of Jacobi
    iterations (update order Q1, fabs convergence test Q3), same amplitudes."""
    case, st = staged
    r = ref.post_hf(case, maxiter=60, conv=1e-12)
    assert abs(r["e_mp2"] - st["e_mp2"]) < 1e-14
    assert r["iters"] == st["iters"]
    assert abs(r["e_ccsd"] - st["e_ccsd"]) < 1e-13
    close(r["T2_mp2"], st["T2_mp2"])
    close(r["T1"], st["T1"], 1e-12)
    close(r["T2"], st["T2"], 1e-12)
    assert abs(r["e_t"] - st["e_t"]) <= 1e-13 * max(1.0, abs(st["e_t"]))


# ---- committed fixture: the reference's results on the real-molecule configs ------------------

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_energies.json")
MOLECULES = {
    "water_sto3g_rhf": dict(xyz="water.xyz", basis="sto-3g"),
    "water_ccpvdz_rhf": dict(xyz="water.xyz", basis="cc-pvdz"),
    "nh2_ccpvdz_uhf": dict(xyz="nh2.xyz", basis="cc-pvdz", uhf=True, multiplicity=2),
}


def test_reference_binary_reproduces_fixture_config1():
    """Re-run the reference program on its own shipped case (water/STO-3G) and compare with the
    committed fixture: guards the fixture against drift of the shims."""
    g = json.load(open(GOLDEN))["water_sto3g_rhf"]
    r = ref.molecule(os.path.join(ROOT, "geom", "water.xyz"), "sto-3g")
    for k in ("e_scf", "e_mp2", "e_ccsd", "e_t"):
        assert abs(r[k] - g[k]) < 1e-12, k
    assert r["cc_iters"] == g["cc_iters"] == 36
    # Crawford's published values for this geometry (the source cc.cpp:453 cites), ~1e-10 (unit constants)
    assert abs(r["e_mp2"] - (-0.049149636120)) < 1e-10 and abs(r["e_ccsd"] - (-0.070680088376)) < 1e-10
    assert abs(r["e_t"] - (-0.000099877272)) < 1e-11


@pytest.mark.parametrize("name", ["water_sto3g_rhf", "water_ccpvdz_rhf"])
def test_oracle_on_host_substrate_matches_reference_fixture(name):
    """Oracle restatement fed by the PRODUCT's host SCF substrate vs the reference program's own SCF +
    post-HF chain: ties the boundary inputs the GPU tests use to what the reference would have seen."""
    g = json.load(open(GOLDEN))[name]
    case = molecule_case(**MOLECULES[name])
    assert (case["n"], case["no"], case["nv"]) == (g["n"], g["no"], g["nv"])
    assert abs(case["e_scf"] - g["e_scf"]) < 1e-10
    st = oracle_pipeline(case, maxiter=200, conv=1e-12, with_t=(name == "water_sto3g_rhf"))
    assert abs(st["e_mp2"] - g["e_mp2"]) < 1e-10 and abs(st["e_ccsd"] - g["e_ccsd"]) < 1e-10
    assert st["iters"] == g["cc_iters"]
    if "e_t" in st:
        assert abs(st["e_t"] - g["e_t"]) < 1e-10


def test_host_uhf_scf_matches_reference_scf():
    """The product's host SCF substrate (qc-p2_b200/host/hf.cpp) vs the reference's UHF (hf.cpp:462-579, compiled) on
    NH2/cc-pVDZ: same SCF energy, orbital energies and -- up to the arbitrary sign of each MO -- coefficients."""
    g = json.load(open(GOLDEN))["nh2_ccpvdz_uhf"]
    b = np.load(os.path.join(ROOT, "tests", "golden", "reference_nh2_boundary.npz"))
    case = molecule_case(**MOLECULES["nh2_ccpvdz_uhf"])
    assert (case["n"], case["no"], case["nv"]) == (g["n"], g["no"], g["nv"])
    assert abs(case["e_scf"] - g["e_scf"]) < 1e-10
    assert np.max(np.abs(case["eps_a"] - b["eps_a"])) < 1e-9 and np.max(np.abs(case["eps_b"] - b["eps_b"])) < 1e-9
    for mine, theirs in ((case["Ca"], b["Ca"]), (case["Cb"], b["Cb"])):
        assert np.max(np.abs(np.abs(mine) - np.abs(theirs))) < 1e-7


def test_reference_vs_oracle_over_random_shapes():
    """Property test (hypothesis): for random small (n, occupation, seed), RHF and as-written UHF, five Jacobi
    iterations of the reference's own chain and of the oracle restatement give the same MP2 / CCSD / (T) numbers."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=20, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(n=st.integers(2, 5), frac=st.floats(0.2, 0.8), seed=st.integers(0, 10 ** 6), uhf=st.booleans())
    def check(n, frac, seed, uhf):
        nd = max(1, min(n - 1, int(round(frac * n))))
        if uhf:
            nb = max(1, nd - 1) if nd > 1 else 1
            case = uhf_case(n, nd, nb, seed) if nd + nb < 2 * n else rhf_case(n, nd, seed)
        else:
            case = rhf_case(n, nd, seed)
        r = ref.post_hf(case, maxiter=5, conv=1e-14)
        o = oracle_pipeline(case, maxiter=5, conv=1e-14)
        scale = max(1.0, abs(o["e_ccsd"]), abs(o["e_t"]))
        assert abs(r["e_mp2"] - o["e_mp2"]) <= 1e-13 * max(1.0, abs(o["e_mp2"]))
        # the reference reports the last iteration number it printed (cc.cpp:437): maxiter - 1 when the cap is hit
        assert r["iters"] == min(o["iters"], 4)
        assert abs(r["e_ccsd"] - o["e_ccsd"]) <= 1e-12 * scale and abs(r["e_t"] - o["e_t"]) <= 1e-12 * scale

    check()
