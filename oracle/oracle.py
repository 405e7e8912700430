"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/liboracle.so
(the literal C++ restatement in oracle.cpp).  Pinned against the compiled reference sources
(oracle/ref.py, tests/test_ref_pins_oracle.py); see the oracle.cpp header.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product path (qc-p2_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
INT_NAMES = ["oooo", "ovoo", "oovo", "ooov", "ovov", "ovvo", "oovv", "vovv", "ovvv", "vvvo", "vvvv"]
INTER_NAMES = ["F_ae", "F_mi", "F_me", "W_mnij", "W_mbej", "W_abef"]

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _ptr_array(arrs):
    arr = (C.POINTER(C.c_double) * len(arrs))()
    for k, a in enumerate(arrs):
        arr[k] = _p(a) if a is not None else None
    return arr


def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle call failed")


def _shape(spec, o, v):
    return tuple(o if c == "o" else v for c in spec)


def inter_shapes(o, v):
    return [(v, v), (o, o), (o, v), (o, o, o, o), (o, v, v, o), (v, v, v, v)]


def contract(alpha, A, ia, B, ib, beta, Cin, ic):
    A, B, Cout = _c(A), _c(B), _c(Cin).copy()
    ext = lambda t: (C.c_long * t.ndim)(*t.shape)
    _chk(lib().orc_contract(C.c_double(alpha), _p(A), ext(A), ia.encode(), _p(B), ext(B), ib.encode(),
                            C.c_double(beta), _p(Cout), ext(Cout), ic.encode()))
    return Cout


def transform_ao_mo(ao, C1, C2):
    ao, C1, C2 = _c(ao), _c(C1), _c(C2)
    n = ao.shape[0]
    out = np.empty((n,) * 4)
    _chk(lib().orc_transform_ao_mo(_p(ao), _p(C1), _p(C2), n, _p(out)))
    return out


def transform_to_so(aa, bb, ab):
    aa, bb, ab = _c(aa), _c(bb), _c(ab)
    n = aa.shape[0]
    out = np.empty((2 * n,) * 4)
    _chk(lib().orc_transform_to_so(_p(aa), _p(bb), _p(ab), n, _p(out)))
    return out


def slice_ints(so, no, nv, spec):
    so = _c(so)
    out = np.empty(_shape(spec, no, nv))
    _chk(lib().orc_slice_ints(_p(so), so.shape[0], no, nv, spec.encode(), _p(out)))
    return out


def get_integrals(so, no, nv):
    return {k: slice_ints(so, no, nv, k) for k in INT_NAMES}


def make_so_moes(ea, eb):
    ea, eb = _c(ea), _c(eb)
    n = ea.shape[0]
    out = np.empty((2 * n, 2 * n))
    _chk(lib().orc_make_so_moes(_p(ea), _p(eb), n, _p(out)))
    return out


def make_denom(F_spin, no, nv):
    F_spin = _c(F_spin)
    out = np.empty((no, no, nv, nv))
    _chk(lib().orc_make_denom(_p(F_spin), F_spin.shape[0], no, nv, _p(out)))
    return out


def run_mp2(oovv, denom):
    oovv, denom = _c(oovv), _c(denom)
    no, nv = oovv.shape[0], oovv.shape[2]
    T = np.empty_like(oovv)
    E = C.c_double()
    _chk(lib().orc_run_mp2(_p(oovv), _p(denom), no, nv, _p(T), C.byref(E)))
    return T, E.value


def make_fock(e1, e2, n1, n2):
    e1, e2 = _c(e1), _c(e2)
    out = np.empty((n2 - n1, n2 - n1))
    _chk(lib().orc_make_fock(_p(e1), _p(e2), n1, n2, _p(out)))
    return out


def make_tau(Ts, Td, bar=False):
    Ts, Td = _c(Ts), _c(Td)
    out = np.empty_like(Td)
    _chk(lib().orc_make_tau(_p(Ts), _p(Td), Ts.shape[0], Ts.shape[1], int(bar), _p(out)))
    return out


def multiply_Ts(Ts):
    Ts = _c(Ts)
    o, v = Ts.shape
    out = np.empty((o, v, o, v))
    _chk(lib().orc_multiply_Ts(_p(Ts), o, v, _p(out)))
    return out


def _ints_list(I):
    return [_c(I.get(k)) for k in INT_NAMES]


def update_intermediates(Ts, Td, I, foo, fov, fvv):
    Ts, Td, foo, fov, fvv = _c(Ts), _c(Td), _c(foo), _c(fov), _c(fvv)
    o, v = Ts.shape
    ints = _ints_list(I)
    outs = [np.empty(s) for s in inter_shapes(o, v)]
    _chk(lib().orc_update_intermediates(_p(Ts), _p(Td), _ptr_array(ints), _p(foo), _p(fov), _p(fvv), o, v,
                                        _ptr_array(outs)))
    return dict(zip(INTER_NAMES, outs))


def make_T1(Ts, Td, I, X, foo, fov, fvv):
    Ts, Td, foo, fov, fvv = _c(Ts), _c(Td), _c(foo), _c(fov), _c(fvv)
    o, v = Ts.shape
    ints, inter = _ints_list(I), [_c(X[k]) for k in INTER_NAMES]
    out = np.empty((o, v))
    _chk(lib().orc_make_T1(_p(Ts), _p(Td), _ptr_array(ints), _ptr_array(inter), _p(foo), _p(fov), _p(fvv), o, v,
                           _p(out)))
    return out


def make_T2(Ts, Td, I, X, foo, fvv):
    Ts, Td, foo, fvv = _c(Ts), _c(Td), _c(foo), _c(fvv)
    o, v = Ts.shape
    ints, inter = _ints_list(I), [_c(X[k]) for k in INTER_NAMES]
    out = np.empty((o, o, v, v))
    _chk(lib().orc_make_T2(_p(Ts), _p(Td), _ptr_array(ints), _ptr_array(inter), _p(foo), _p(fvv), o, v, _p(out)))
    return out


def ccsd_energy(Ts, Td, oovv, fov):
    Ts, Td, oovv, fov = _c(Ts), _c(Td), _c(oovv), _c(fov)
    E = C.c_double()
    _chk(lib().orc_ccsd_energy(_p(Ts), _p(Td), _p(oovv), _p(fov), Ts.shape[0], Ts.shape[1], C.byref(E)))
    return E.value


def ccsd(I, foo, fov, fvv, T2_guess, maxiter, conv):
    foo, fov, fvv, T2_guess = _c(foo), _c(fov), _c(fvv), _c(T2_guess)
    o, v = fov.shape
    ints = _ints_list(I)
    T1, T2 = np.empty((o, v)), np.empty((o, o, v, v))
    E, iters = C.c_double(), C.c_int()
    hist = np.zeros(maxiter)
    _chk(lib().orc_ccsd(_ptr_array(ints), _p(foo), _p(fov), _p(fvv), _p(T2_guess), o, v, maxiter, C.c_double(conv),
                        _p(T1), _p(T2), C.byref(E), C.byref(iters), _p(hist)))
    n = min(iters.value + 1, maxiter)
    return T1, T2, E.value, iters.value, hist[:n]


def ccsd_t(T1, T2, oovv, vovv, ovoo, foo, fvv):
    T1, T2, oovv, vovv, ovoo, foo, fvv = map(_c, (T1, T2, oovv, vovv, ovoo, foo, fvv))
    E = C.c_double()
    _chk(lib().orc_ccsd_t(_p(T1), _p(T2), _p(oovv), _p(vovv), _p(ovoo), _p(foo), _p(fvv), T1.shape[0], T1.shape[1],
                          C.byref(E)))
    return E.value


def ccsd_t_blocked(T1, T2, oovv, vovv, ovoo, eo, ev, triples, weight=1.0):
    T1, T2, oovv, vovv, ovoo, eo, ev = map(_c, (T1, T2, oovv, vovv, ovoo, eo, ev))
    tr = np.ascontiguousarray(triples, dtype=np.int32).reshape(-1, 3)
    E = C.c_double()
    _chk(lib().orc_ccsd_t_blocked(_p(T1), _p(T2), _p(oovv), _p(vovv), _p(ovoo), _p(eo), _p(ev), T1.shape[0],
                                  T1.shape[1], tr.ctypes.data_as(C.POINTER(C.c_int)), C.c_long(tr.shape[0]),
                                  C.c_double(weight), C.byref(E)))
    return E.value


def t_energy_packed(Xp, T1, oovv, eo, ev, i, j, k):
    """sum_{a<b<c} W (W+V)/D of one sorted triple from the b<c-packed X (OpenMP; CPU baseline)."""
    Xp, T1, eo, ev = map(_c, (Xp, T1, eo, ev))
    if not (oovv.flags["C_CONTIGUOUS"] and oovv.dtype == np.float64):
        oovv = _c(oovv)
    E = C.c_double()
    _chk(lib().orc_t_energy_packed(_p(Xp), _p(T1), _p(oovv), _p(eo), _p(ev), T1.shape[0], T1.shape[1], int(i), int(j), int(k),
                                   C.byref(E)))
    return E.value
