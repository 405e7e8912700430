"""Known-answer tests (CPU): the only external anchors of the oracle.

The reference has no tests or golden vectors.  What exists outside the tree: the Crawford-group
ProgrammingProjects reference output for exactly geom/water.xyz / STO-3G, the source the
reference cites at cc.cpp:453.  The numbers below are quoted in SURVEY.md section 8c / BASELINE.md
"from memory"; agreement is limited to ~1e-9 by the Angstrom->bohr constant (E_nuc differs by
1.6e-9).  The boundary inputs come from the host SCF substrate (qc-p2_b200/host), the energies
from the literal C++ oracle -- so this pins the oracle's MP2, CCSD and (T) to 1e-10."""
import numpy as np
import pytest
from helpers import fock_blocks, load_hostscf, molecule_case, oracle_pipeline

CRAWFORD = dict(e_nuc=8.002367061810450, e_scf=-74.942079928192, e_mp2=-0.049149636120, e_ccsd=-0.070680088376,
                e_t=-0.000099877272)


@pytest.fixture(scope="module")
def water_sto3g():
    case = molecule_case("water.xyz", "sto-3g")
    return case, oracle_pipeline(case, maxiter=200, conv=1e-12)


def test_water_sto3g_scf(water_sto3g):
    case, _ = water_sto3g
    assert (case["n"], case["no"], case["nv"]) == (7, 10, 4)
    assert abs(case["e_nuc"] - CRAWFORD["e_nuc"]) < 5e-9
    assert abs(case["e_scf"] - CRAWFORD["e_scf"]) < 1e-9


def test_water_sto3g_mp2_ccsd_t_match_crawford(water_sto3g):
    _, r = water_sto3g
    assert abs(r["e_mp2"] - CRAWFORD["e_mp2"]) < 1e-10
    assert abs(r["e_ccsd"] - CRAWFORD["e_ccsd"]) < 1e-10
    assert abs(r["e_t"] - CRAWFORD["e_t"]) < 1e-11


def test_water_sto3g_cross_oracle(water_sto3g):
    from oracle import np_oracle as npo
    case, r = water_sto3g
    In = npo.get_integrals(r["so"], case["no"])
    _, _, hist = npo.ccsd(In, r["foo"], r["fov"], r["fvv"], r["T2_mp2"], 200, 1e-12)
    assert abs(hist[-1] - r["e_ccsd"]) < 1e-12
    eo, ev = np.diag(r["foo"]).copy(), np.diag(r["fvv"]).copy()
    assert abs(npo.ccsd_t_blocked(r["T1"], r["T2"], In["oovv"], In["vovv"], In["ovoo"], eo, ev, True) - r["e_t"]) < 1e-14


def test_builtin_basis_contractions_are_normalised():
    hs = load_hostscf()
    for basis, elements in (("sto-3g", (1, 6, 7, 8)), ("cc-pvdz", (1, 6, 7, 8)), ("cc-pvtz", (1, 6))):
        for Z in elements:
            for l, norm in hs.basis_norms(basis, Z):
                assert abs(norm - 1.0) < 2e-4, (basis, Z, l, norm)


def test_water_ccpvdz_scf_exact_integrals(tmp_path):
    """H2O, R = 0.96 A, 104.5 deg, cc-pVDZ, conventional (non-density-fitted) RHF: -76.02665366 Eh
    (Psi4 test-suite value, quoted from memory; the density-fitted tutorial number -76.02663273
    differs by the DF error).  Exercises pure d functions and the typed-in cc-pVDZ data."""
    import math
    R, th = 0.96, math.radians(104.5)
    xyz = tmp_path / "h2o.xyz"
    xyz.write_text("3\n\nO 0 0 0\nH %.12f 0 0\nH %.12f %.12f 0\n" % (R, R * math.cos(th), R * math.sin(th)))
    r = load_hostscf().scf(str(xyz), "cc-pvdz")
    assert r["n"] == 24
    assert abs(r["e_scf"] - (-76.02665366)) < 2e-7


def test_scf_energy_is_rotation_invariant(tmp_path):
    """d-function integrals: rotating the molecule must not change the energy."""
    hs = load_hostscf()
    base = np.array([[0.0, -0.0757918381, 0.0], [0.8668117664, 0.6014357359, 0.0], [-0.8668117664, 0.6014357359, 0.0]])
    a, b = 0.7, 1.9
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    es = []
    for rot in (np.eye(3), Rz @ Rx):
        xyz = tmp_path / "w.xyz"
        pts = base @ rot.T + (0.0 if rot is None else 0.3)
        xyz.write_text("3\n\n" + "\n".join("%s %.12f %.12f %.12f" % (s, *p) for s, p in zip("OHH", pts)) + "\n")
        es.append(hs.scf(str(xyz), "cc-pvdz", conv=1e-11)["e_scf"])
    assert abs(es[0] - es[1]) < 1e-9


def test_config3_oracle_golden_equals_the_reference_s_own_mp2_and_ccsd():
    """CH4/cc-pVTZ: the committed CPU-oracle numbers (tests/golden/ch4_ccpvtz_oracle.json) against the reference's OWN MP2() /
    CCSD() run on the same molecule (tests/golden/reference_energies.json, entry written by make_reference_golden_ch4.py --
    37 minutes of the reference's unmodified sources on the shim's OpenMP GEMM).  CCSD_T() cannot hold this config."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    orc_g = json.load(open(os.path.join(root, "tests", "golden", "ch4_ccpvtz_oracle.json")))
    ref_g = json.load(open(os.path.join(root, "tests", "golden", "reference_energies.json")))["ch4_ccpvtz_rhf"]
    assert (ref_g["n"], ref_g["no"], ref_g["nv"]) == (86, 10, 162)
    assert abs(orc_g["e_scf"] - ref_g["e_scf"]) < 1e-9
    assert abs(orc_g["e_mp2"] - ref_g["e_mp2"]) < 1e-10 and abs(orc_g["e_ccsd"] - ref_g["e_ccsd"]) < 1e-10
    assert orc_g["iters"] == ref_g["cc_iters"] == 29
