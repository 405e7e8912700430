"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/_ref/libqcp2_ref.so, i.e.
the reference's OWN sources (/root/reference/src/{integrals,mp2,cc,hf,general}.cpp, unmodified)
compiled against the API shims in oracle/ref_shim/include (BTAS, Eigen, Libint2, nlohmann-json
stand-ins -- the real libraries are absent from this image).  It is what pins the restatement in
oracle.cpp: tests/test_ref_pins_oracle.py compares the two function by function.

Only tests/, __graft_entry__ and bench.py's reference / cpu_baseline legs may import this module.
The library is built here (where /root/reference exists) by `make -C oracle/ref_shim`; the GPU box
only sees the prebuilt files, which travel with the snapshot.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libqcp2_ref.so")
BIN = os.path.join(HERE, "_ref", "P2_ref")
REFERENCE = "/root/reference"
INT_NAMES = ["oooo", "ovoo", "oovo", "ooov", "ovov", "ovvo", "oovv", "vovv", "ovvv", "vvvo", "vvvv"]
INTER_NAMES = ["F_ae", "F_mi", "F_me", "W_mnij", "W_mbej", "W_abef"]

_lib = None


def can_build():
    return os.path.isdir(os.path.join(REFERENCE, "src"))


def build():
    """Compile the reference from where it lies (needs /root/reference; a no-op when up to date)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "ref_shim")])


def available():
    if can_build():
        return True
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if can_build():
            build()
        if not os.path.exists(LIB):
            raise RuntimeError("oracle/_ref/libqcp2_ref.so is missing and /root/reference is not here to build it")
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
        raise RuntimeError("reference call failed")


def _shape(spec, o, v):
    return tuple(o if c == "o" else v for c in spec)


def inter_shapes(o, v):
    return [(v, v), (o, o), (o, v), (o, o, o, o), (o, v, v, o), (v, v, v, v)]


def shim_contract(alpha, A, ia, B, ib, beta, Cin, ic):
    """The shim's btas::contract (oracle/ref_shim/include/btas/btas.h) on label strings -- for testing the shim."""
    A, B, Cout = _c(A), _c(B), _c(Cin).copy()
    ext = lambda t: (C.c_long * t.ndim)(*t.shape)
    _chk(lib().ref_shim_contract(C.c_double(alpha), _p(A), ext(A), ia.encode(), _p(B), ext(B), ib.encode(),
                                 C.c_double(beta), _p(Cout), ext(Cout), ic.encode()))
    return Cout


def transform_ao_mo(ao, C1, C2):
    """integrals.cpp:95-134"""
    ao, C1, C2 = _c(ao), _c(C1), _c(C2)
    n = ao.shape[0]
    out = np.empty((n,) * 4)
    _chk(lib().ref_transform_ao_mo(_p(ao), _p(C1), _p(C2), n, _p(out)))
    return out


def transform_to_so(aa, bb, ab):
    """integrals.cpp:136-167"""
    aa, bb, ab = _c(aa), _c(bb), _c(ab)
    n = aa.shape[0]
    out = np.empty((2 * n,) * 4)
    _chk(lib().ref_transform_to_so(_p(aa), _p(bb), _p(ab), n, _p(out)))
    return out


def slice_ints(so, no, nv, spec):
    """integrals.cpp:184-207"""
    so = _c(so)
    out = np.empty(_shape(spec, no, nv))
    _chk(lib().ref_slice_ints(_p(so), so.shape[0], no, nv, spec.encode(), _p(out)))
    return out


def get_integrals(so, no, nv):
    return {k: slice_ints(so, no, nv, k) for k in INT_NAMES}


def make_so_moes(ea, eb):
    """mp2.cpp:11-21"""
    ea, eb = _c(ea), _c(eb)
    n = ea.shape[0]
    out = np.empty((2 * n, 2 * n))
    _chk(lib().ref_make_so_moes(_p(ea), _p(eb), n, _p(out)))
    return out


def make_denom(F_spin, no, nv):
    """integrals.cpp:169-181"""
    F_spin = _c(F_spin)
    out = np.empty((no, no, nv, nv))
    _chk(lib().ref_make_denom(_p(F_spin), F_spin.shape[0], no, nv, _p(out)))
    return out


def run_mp2(oovv, denom):
    """mp2.cpp:23-42"""
    oovv, denom = _c(oovv), _c(denom)
    no, nv = oovv.shape[0], oovv.shape[2]
    T = np.empty_like(oovv)
    E = C.c_double()
    _chk(lib().ref_run_mp2(_p(oovv), _p(denom), no, nv, _p(T), C.byref(E)))
    return T, E.value


def make_fock(e1, e2, n1, n2):
    """cc.cpp:25-34"""
    e1, e2 = _c(e1), _c(e2)
    out = np.empty((n2 - n1, n2 - n1))
    _chk(lib().ref_make_fock(_p(e1), _p(e2), e1.shape[0], n1, n2, _p(out)))
    return out


def make_D(foo, fvv, kind):
    """cc.cpp:73-123; kind 0: D_ia, 1: D_ijab, 2: D_triples"""
    foo, fvv = _c(foo), _c(fvv)
    o, v = foo.shape[0], fvv.shape[0]
    out = np.empty([(o, v), (o, o, v, v), (o, o, o, v, v, v)][kind])
    _chk(lib().ref_make_D(_p(foo), _p(fvv), o, v, kind, _p(out)))
    return out


def make_tau(Ts, Td, bar=False):
    """cc.cpp:125-156"""
    Ts, Td = _c(Ts), _c(Td)
    out = np.empty_like(Td)
    _chk(lib().ref_make_tau(_p(Ts), _p(Td), Ts.shape[0], Ts.shape[1], int(bar), _p(out)))
    return out


def multiply_Ts(Ts):
    """cc.cpp:158-172"""
    Ts = _c(Ts)
    o, v = Ts.shape
    out = np.empty((o, v, o, v))
    _chk(lib().ref_multiply_Ts(_p(Ts), o, v, _p(out)))
    return out


def _ints_list(I):
    return [_c(I.get(k)) for k in INT_NAMES]


def update_intermediates(Ts, Td, I, foo, fov, fvv):
    """cc.cpp:175-255"""
    Ts, Td, foo, fov, fvv = _c(Ts), _c(Td), _c(foo), _c(fov), _c(fvv)
    o, v = Ts.shape
    ints = _ints_list(I)
    outs = [np.empty(s) for s in inter_shapes(o, v)]
    _chk(lib().ref_update_intermediates(_p(Ts), _p(Td), _ptr_array(ints), _p(foo), _p(fov), _p(fvv), o, v,
                                        _ptr_array(outs)))
    return dict(zip(INTER_NAMES, outs))


def make_T1(Ts, Td, I, X, foo, fov, fvv):
    """cc.cpp:257-292"""
    Ts, Td, foo, fov, fvv = _c(Ts), _c(Td), _c(foo), _c(fov), _c(fvv)
    o, v = Ts.shape
    ints, inter = _ints_list(I), [_c(X[k]) for k in INTER_NAMES]
    out = np.empty((o, v))
    _chk(lib().ref_make_T1(_p(Ts), _p(Td), _ptr_array(ints), _ptr_array(inter), _p(foo), _p(fov), _p(fvv), o, v,
                           _p(out)))
    return out


def make_T2(Ts, Td, I, X, foo, fvv):
    """cc.cpp:294-387"""
    Ts, Td, foo, fvv = _c(Ts), _c(Td), _c(foo), _c(fvv)
    o, v = Ts.shape
    ints, inter = _ints_list(I), [_c(X[k]) for k in INTER_NAMES]
    out = np.empty((o, o, v, v))
    _chk(lib().ref_make_T2(_p(Ts), _p(Td), _ptr_array(ints), _ptr_array(inter), _p(foo), _p(fvv), o, v, _p(out)))
    return out


def ccsd_energy(Ts, Td, oovv, fov):
    """cc.cpp:390-405"""
    Ts, Td, oovv, fov = _c(Ts), _c(Td), _c(oovv), _c(fov)
    E = C.c_double()
    _chk(lib().ref_ccsd_energy(_p(Ts), _p(Td), _p(oovv), _p(fov), Ts.shape[0], Ts.shape[1], C.byref(E)))
    return E.value


def ccsd_t(T1, T2, oovv, vovv, ovoo, foo, fvv):
    """cc.cpp:454-560 (materialises four o^3 v^3 tensors: small shapes only)"""
    T1, T2, oovv, vovv, ovoo, foo, fvv = map(_c, (T1, T2, oovv, vovv, ovoo, foo, fvv))
    E = C.c_double()
    _chk(lib().ref_ccsd_t(_p(T1), _p(T2), _p(oovv), _p(vovv), _p(ovoo), _p(foo), _p(fvv), T1.shape[0], T1.shape[1],
                          C.byref(E)))
    return E.value


def post_hf(case, maxiter=100, conv=1e-12, with_t=True, with_ccsd=True):
    """The reference's MP2 -> CCSD -> (T) chain (mp2.cpp:51-83 minus the Libint call, cc.cpp:409-451,
    main.cpp:84-85) on a tests/helpers rhf_case / uhf_case dict."""
    uhf = "Ca" in case
    n, no, nv = case["n"], case["no"], case["nv"]
    ao = _c(case["ao"])
    Ca, Cb = (_c(case["Ca"]), _c(case["Cb"])) if uhf else (_c(case["C"]), _c(case["C"]))
    ea, eb = (_c(case["eps_a"]), _c(case["eps_b"])) if uhf else (_c(case["eps"]), _c(case["eps"]))
    e_mp2, e_cc, e_t, iters = C.c_double(), C.c_double(), C.c_double(), C.c_int(-1)
    T1, T2, T2g = np.zeros((no, nv)), np.zeros((no, no, nv, nv)), np.zeros((no, no, nv, nv))
    _chk(lib().ref_post_hf(_p(ao), _p(Ca), _p(Cb), _p(ea), _p(eb), n, no, nv, int(uhf), maxiter, C.c_double(conv),
                           int(with_t), C.byref(e_mp2), C.byref(e_cc) if with_ccsd else None, C.byref(e_t),
                           C.byref(iters), _p(T1), _p(T2), _p(T2g)))
    out = dict(e_mp2=e_mp2.value, T2_mp2=T2g)
    if with_ccsd:
        out.update(e_ccsd=e_cc.value, iters=iters.value, T1=T1, T2=T2)
        if with_t:
            out["e_t"] = e_t.value
    return out


def molecule(xyz_path, basis, ref="RHF", type_="CCSD(T)", multiplicity=1, maxiter=200, scf_conv=1e-12, cc_conv=1e-12,
             want_boundary=False):
    """The whole reference program for one molecule (main.cpp:22-88: read_geometry, BasisSet, RHF/UHF,
    MP2, CCSD, CCSD_T), energies at full precision."""
    nao, no, nv, it = C.c_int(), C.c_int(), C.c_int(), C.c_int(-1)
    e = [C.c_double(float("nan")) for _ in range(4)]
    args = lambda ao, Ca, Cb, ea, eb: (xyz_path.encode(), basis.encode(), ref.encode(), type_.encode(), multiplicity,
                                       maxiter, C.c_double(scf_conv), C.c_double(cc_conv), C.byref(nao), C.byref(no),
                                       C.byref(nv), C.byref(e[0]), C.byref(e[1]), C.byref(e[2]), C.byref(e[3]),
                                       C.byref(it), _p(ao), _p(Ca), _p(Cb), _p(ea), _p(eb))
    out = {}
    if want_boundary:
        _chk(lib().ref_nbasis(xyz_path.encode(), basis.encode(), C.byref(nao)))
        n = nao.value
        ao, Ca, Cb, ea, eb = np.empty((n,) * 4), np.empty((n, n)), np.empty((n, n)), np.empty(n), np.empty(n)
        _chk(lib().ref_molecule(*args(ao, Ca, Cb, ea, eb)))
        out.update(ao=ao, Ca=Ca, Cb=Cb, eps_a=ea, eps_b=eb)
    else:
        _chk(lib().ref_molecule(*args(None, None, None, None, None)))
    out.update(n=nao.value, no=no.value, nv=nv.value, e_scf=e[0].value, e_mp2=e[1].value, e_ccsd=e[2].value,
               e_t=e[3].value, cc_iters=it.value)
    return out
