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
