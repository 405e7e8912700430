"""CPU tests: the two oracles against each other and against the committed golden fixtures.
(The reference has no tests and cannot be built here -- see oracle/oracle.cpp header.)"""
import json
import os

import numpy as np
import pytest
from helpers import fock_blocks, oracle_pipeline, rhf_case, uhf_case

from oracle import np_oracle as npo
from oracle import oracle as orc

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "energies.json")))


def build_case(spec):
    if spec["kind"] == "rhf":
        return rhf_case(spec["n"], spec["ndocc"], spec["seed"])
    return uhf_case(spec["n"], spec["na"], spec["nb"], spec["seed"])


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_reproduces_golden(name):
    g = GOLDEN[name]
    r = oracle_pipeline(build_case(g["spec"]))
    assert abs(r["e_mp2"] - g["e_mp2"]) < 1e-13
    assert r["iters"] == g["iters"]
    assert np.max(np.abs(np.array(r["e_hist"]) - np.array(g["e_hist"]))) < 1e-13
    assert abs(r["e_ccsd"] - g["e_ccsd"]) < 1e-13
    assert abs(r["e_t"] - g["e_t"]) < 1e-14
    chk = float(np.sum(r["so"] * np.arange(1, r["so"].size + 1).reshape(r["so"].shape)))
    assert abs(chk - g["so_checksum"]) < 1e-9 * max(1.0, abs(g["so_checksum"]))


@pytest.mark.parametrize("pattern", [
    ("pqrs", "sl", "pqrl"), ("sl", "pqrs", "pqrl"), ("qj", "pqkl", "pjkl"), ("me", "ma", "ae"), ("mf", "mafe", "ae"),
    ("mnaf", "mnef", "ae"), ("nf", "mnef", "me"), ("ijef", "mnef", "mnij"), ("mb", "amef", "abef"),
    ("jnfb", "mnef", "mbej"), ("imae", "mbej", "ijab"), ("iema", "mbje", "ijab"), ("jkae", "eibc", "ijkabc"),
    ("imbc", "majk", "ijkabc")])
def test_contract_matches_einsum(pattern):
    ia, ib, ic = pattern
    rng = np.random.default_rng(7)
    ext = {l: int(e) for l, e in zip("abcefijklmnpqrs", rng.integers(2, 5, 15))}
    A = rng.standard_normal([ext[l] for l in ia])
    B = rng.standard_normal([ext[l] for l in ib])
    C0 = rng.standard_normal([ext[l] for l in ic])
    got = orc.contract(0.7, A, ia, B, ib, -0.3, C0, ic)
    want = -0.3 * C0 + 0.7 * np.einsum(f"{ia},{ib}->{ic}", A, B)
    assert np.max(np.abs(got - want)) < 1e-13


def test_rhf_cross_oracle_full_pipeline():
    case = rhf_case(6, 2, 21)
    r = oracle_pipeline(case)
    no = case["no"]
    mo = npo.transform_ao_mo(case["ao"], case["C"], case["C"])
    assert np.max(np.abs(mo - r["mo_aa"])) < 1e-14
    so = npo.transform_to_so(mo, mo, mo)
    assert np.max(np.abs(so - r["so"])) < 1e-14
    In = npo.get_integrals(so, no)
    eo, ev = np.diag(r["foo"]).copy(), np.diag(r["fvv"]).copy()
    T2, e2 = npo.mp2(In["oovv"], eo, ev)
    assert abs(e2 - r["e_mp2"]) < 1e-14
    t1, t2, hist = npo.ccsd(In, r["foo"], r["fov"], r["fvv"], T2, 100, 1e-12)
    assert len(hist) == len(r["e_hist"])
    assert np.max(np.abs(np.array(hist) - r["e_hist"])) < 1e-13
    assert np.max(np.abs(t2 - r["T2"])) < 1e-13 and np.max(np.abs(t1 - r["T1"])) < 1e-13
    et = npo.ccsd_t_einsum(r["T1"], r["T2"], In["oovv"], In["vovv"], In["ovoo"], eo, ev)
    assert abs(et - r["e_t"]) < 1e-15
    # single-GEMM-per-triple restatement, literal and symmetry-reduced
    assert abs(npo.ccsd_t_blocked(r["T1"], r["T2"], In["oovv"], In["vovv"], In["ovoo"], eo, ev, False) - r["e_t"]) < 1e-15
    assert abs(npo.ccsd_t_blocked(r["T1"], r["T2"], In["oovv"], In["vovv"], In["ovoo"], eo, ev, True) - r["e_t"]) < 1e-15


def test_rhf_so_ints_have_physical_symmetries_and_uhf_as_written_does_not():
    r = oracle_pipeline(rhf_case(4, 2, 2), with_t=False, maxiter=1)
    so = r["so"]
    assert np.max(np.abs(so + so.transpose(1, 0, 2, 3))) < 1e-14
    assert np.max(np.abs(so + so.transpose(0, 1, 3, 2))) < 1e-14
    assert np.max(np.abs(so - so.transpose(2, 3, 0, 1))) < 1e-14
    u = oracle_pipeline(uhf_case(4, 2, 1, 2), with_t=False, maxiter=1)["so"]
    assert np.max(np.abs(u + u.transpose(1, 0, 2, 3))) < 1e-14          # bra antisymmetry survives
    assert np.max(np.abs(u + u.transpose(0, 1, 3, 2))) > 1e-6           # ket antisymmetry is lost


def test_uhf_as_written_literal_vs_blocked_t():
    case = uhf_case(5, 3, 2, 9)
    r = oracle_pipeline(case)
    eo, ev = np.diag(r["foo"]).copy(), np.diag(r["fvv"]).copy()
    I = r["ints"]
    lit = npo.ccsd_t_blocked(r["T1"], r["T2"], I["oovv"], I["vovv"], I["ovoo"], eo, ev, False)
    assert abs(lit - r["e_t"]) < 1e-15
    blk = orc.ccsd_t_blocked(r["T1"], r["T2"], I["oovv"], I["vovv"], I["ovoo"], eo, ev, npo.all_triples(case["no"]))
    assert abs(blk - r["e_t"]) < 1e-15


def test_synthetic_inputs_are_antisymmetric_and_deterministic():
    s = npo.synthetic_t_inputs(5, 7)
    s2 = npo.synthetic_t_inputs(5, 7)
    for k in s:
        assert np.array_equal(s[k], s2[k])
    assert np.array_equal(s["T2"], -s["T2"].transpose(1, 0, 2, 3)) and np.array_equal(s["T2"], -s["T2"].transpose(0, 1, 3, 2))
    assert np.array_equal(s["oovv"], -s["oovv"].transpose(0, 1, 3, 2))
    assert np.array_equal(s["vovv"], -s["vovv"].transpose(0, 1, 3, 2))
    assert np.array_equal(s["ovoo"], -s["ovoo"].transpose(0, 1, 3, 2))
    assert np.all(np.abs(s["T1"]) <= 0.05) and abs(s["eps_o"][0] + 2.0) < 1e-15 and abs(s["eps_v"][-1] - 3.0) < 1e-15
    # known answer of the generator itself (splitmix64 of 0 is 0xE220A8397B1DCDAF)
    assert int(npo.splitmix64(np.uint64(0))) == 0xE220A8397B1DCDAF


def test_synthetic_t_literal_vs_symmetric_vs_packed():
    s = npo.synthetic_t_inputs(4, 6)
    args = (s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"])
    lit = npo.ccsd_t_blocked(*args, False)
    sym = npo.ccsd_t_blocked(*args, True)
    ein = npo.ccsd_t_einsum(*args)
    packed = sum(npo.ccsd_t_triple_packed_np(*args, int(i), int(j), int(k)) for i, j, k in npo.sorted_triples(4))
    packed_fast = sum(npo.ccsd_t_triple_packed_np(*args, int(i), int(j), int(k), fast=True) for i, j, k in npo.sorted_triples(4))
    assert abs(packed_fast - packed) < 1e-14 * max(1.0, abs(packed))
    pk = npo.PackedTriples(*args)
    assert abs(sum(pk.triple(int(i), int(j), int(k)) for i, j, k in npo.sorted_triples(4)) - packed) < 1e-14 * max(1.0, abs(packed))
    foo, fvv = np.diag(s["eps_o"]), np.diag(s["eps_v"])
    cpp = orc.ccsd_t(s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], foo, fvv)
    for x in (sym, ein, packed, cpp):
        assert abs(x - lit) < 1e-13 * max(1.0, abs(lit))


def test_empty_virtual_or_tiny_spaces():
    # o=2, v=2: no i<j<k triple exists -> E(T) == 0 in every formulation
    s = npo.synthetic_t_inputs(2, 2)
    args = (s["T1"], s["T2"], s["oovv"], s["vovv"], s["ovoo"], s["eps_o"], s["eps_v"])
    assert npo.ccsd_t_blocked(*args, True) == 0.0
    assert abs(npo.ccsd_t_blocked(*args, False)) < 1e-18
