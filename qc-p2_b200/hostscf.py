"""hostscf.py -- ctypes front end of lib/libqcp2_host.so, the HOST substrate (Gaussian
integrals + RHF/UHF, qc-p2_b200/host/).  It produces the hot-path boundary inputs of the
real-molecule configs: AO integrals (pq|rs), MO coefficients [AO][MO] and orbital energies."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libqcp2_host.so")
_lib = None
PD = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "host"), "../lib/libqcp2_host.so"])
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(PD)


def scf(xyz_path, basis, uhf=False, multiplicity=1, maxiter=200, conv=1e-12, quiet=True):
    """Runs the reference-style Roothaan SCF; returns a dict with ao, Ca, Cb, eps_a, eps_b, S, n, no, nv,
    e_scf, e_nuc, iters (for RHF Ca is Cb and eps_a is eps_b)."""
    n, no, nv, it = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    es, en = C.c_double(), C.c_double()
    args = (xyz_path.encode(), basis.encode(), int(uhf), multiplicity, maxiter, C.c_double(conv))
    rc = lib().qcp2_host_scf(*args, C.byref(n), C.byref(no), C.byref(nv), C.byref(es), C.byref(en), C.byref(it), None, None,
                             None, None, None, None, 1)
    if rc:
        raise RuntimeError("host SCF failed (geometry/basis)")
    N = n.value
    ao, Ca, Cb, S = np.empty((N,) * 4), np.empty((N, N)), np.empty((N, N)), np.empty((N, N))
    ea, eb = np.empty(N), np.empty(N)
    rc = lib().qcp2_host_scf(*args, C.byref(n), C.byref(no), C.byref(nv), C.byref(es), C.byref(en), C.byref(it), _p(ao),
                             _p(Ca), _p(Cb), _p(ea), _p(eb), _p(S), int(quiet))
    if rc:
        raise RuntimeError("host SCF failed")
    return dict(ao=ao, Ca=Ca, Cb=Cb, eps_a=ea, eps_b=eb, S=S, n=N, no=no.value, nv=nv.value, e_scf=es.value,
                e_nuc=en.value, iters=it.value, uhf=bool(uhf))


def basis_norms(basis, Z):
    norms = np.zeros(64)
    ls = (C.c_int * 64)()
    k = lib().qcp2_host_basis_norms(basis.encode(), Z, _p(norms), ls, 64)
    if k < 0:
        raise RuntimeError("basis lookup failed")
    return [(ls[i], norms[i]) for i in range(k)]
