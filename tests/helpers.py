"""Shared test helpers: seeded pseudo-molecules at the hot-path boundary (ao_ints, C, eps)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def fake_ao(n, seed, scale=0.08):
    """Random AO two-electron tensor with the 8-fold permutational symmetry of real (pq|rs)."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n, n, n))
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    return np.ascontiguousarray(g * (scale / 8.0))


def fake_orbitals(n, nocc, seed, gap=1.0):
    """Random orthogonal MO coefficients [AO][MO] and sorted orbital energies with an o/v gap."""
    rng = np.random.default_rng(seed + 1000)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    occ = np.sort(-gap / 2 - 1.5 * rng.random(nocc))
    vir = np.sort(gap / 2 + 2.5 * rng.random(n - nocc))
    return np.ascontiguousarray(q), np.concatenate([occ, vir])


def rhf_case(n, ndocc, seed):
    ao = fake_ao(n, seed)
    C, eps = fake_orbitals(n, ndocc, seed)
    return dict(ao=ao, C=C, eps=eps, n=n, no=2 * ndocc, nv=2 * (n - ndocc))


def uhf_case(n, nalpha, nbeta, seed):
    ao = fake_ao(n, seed)
    Ca, ea = fake_orbitals(n, nalpha, seed)
    Cb, eb = fake_orbitals(n, nbeta, seed + 7)
    no = nalpha + nbeta
    return dict(ao=ao, Ca=Ca, Cb=Cb, eps_a=ea, eps_b=eb, n=n, no=no, nv=2 * n - no)


def fock_blocks(eps_a, eps_b, no):
    """Diagonal spin-orbital Fock blocks as make_moe_tensors builds them (cc.cpp:36-70): alpha on
    even, beta on odd positions; F_ia is zero (RHF) / the only defined reading for UHF (Q4)."""
    n2 = 2 * len(eps_a)
    d = np.empty(n2)
    d[0::2], d[1::2] = eps_a, eps_b
    return np.diag(d[:no]).copy(), np.zeros((no, n2 - no)), np.diag(d[no:]).copy()


def oracle_pipeline(case, maxiter=100, conv=1e-12, with_t=True):
    """MP2 -> CCSD -> (T) with the literal C++ oracle on a helpers.rhf_case / uhf_case."""
    from oracle import oracle as orc
    uhf = "Ca" in case
    no, nv = case["no"], case["nv"]
    if uhf:
        aa = orc.transform_ao_mo(case["ao"], case["Ca"], case["Ca"])
        bb = orc.transform_ao_mo(case["ao"], case["Cb"], case["Cb"])
        ab = orc.transform_ao_mo(case["ao"], case["Ca"], case["Cb"])
        ea, eb = case["eps_a"], case["eps_b"]
    else:
        aa = bb = ab = orc.transform_ao_mo(case["ao"], case["C"], case["C"])
        ea = eb = case["eps"]
    so = orc.transform_to_so(aa, bb, ab)
    F = orc.make_so_moes(ea, eb)
    I = orc.get_integrals(so, no, nv)
    T2_mp2, e_mp2 = orc.run_mp2(I["oovv"], orc.make_denom(F, no, nv))
    foo, fov, fvv = fock_blocks(ea, eb, no)
    T1, T2, e_cc, iters, hist = orc.ccsd(I, foo, fov, fvv, T2_mp2, maxiter, conv)
    out = dict(mo_aa=aa, mo_bb=bb, mo_ab=ab, so=so, F=F, ints=I, T2_mp2=T2_mp2, e_mp2=e_mp2, foo=foo, fov=fov, fvv=fvv,
               T1=T1, T2=T2, e_ccsd=e_cc, iters=iters, e_hist=hist, eps_a=ea, eps_b=eb)
    if with_t:
        out["e_t"] = orc.ccsd_t(T1, T2, I["oovv"], I["vovv"], I["ovoo"], foo, fvv)
    return out


def load_hostscf():
    import importlib.util
    spec = importlib.util.spec_from_file_location("qcp2_hostscf", os.path.join(ROOT, "qc-p2_b200", "hostscf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def molecule_case(xyz, basis, uhf=False, multiplicity=1, conv=1e-12, maxiter=200):
    """Boundary inputs of a real-molecule config from the host SCF substrate, in the dict
    layout of rhf_case / uhf_case."""
    r = load_hostscf().scf(os.path.join(ROOT, "geom", xyz), basis, uhf=uhf, multiplicity=multiplicity, conv=conv,
                           maxiter=maxiter)
    if uhf:
        return dict(ao=r["ao"], Ca=r["Ca"], Cb=r["Cb"], eps_a=r["eps_a"], eps_b=r["eps_b"], n=r["n"], no=r["no"],
                    nv=r["nv"], e_scf=r["e_scf"], e_nuc=r["e_nuc"])
    return dict(ao=r["ao"], C=r["Ca"], eps=r["eps_a"], n=r["n"], no=r["no"], nv=r["nv"], e_scf=r["e_scf"],
                e_nuc=r["e_nuc"])
