"""qcp2.py -- Python host mirror of the reference's post-HF interface (integrals.h, mp2.h,
cc.h) on top of the C ABI in include/qcp2_b200.h (lib/libqcp2_b200.so).

Function names and argument meaning follow the reference (transform_ao_mo, transform_to_so,
slice_ints, make_denom, make_so_moes, run_mp2, MP2, make_fock, make_tau, make_tau_bar,
multiply_Ts, make_D_ia, make_D_ijab, make_D_triples, update_intermediates, make_T1, make_T2,
ccsd_energy, CCSD, CCSD_T).  Tensors are numpy float64 arrays (host pointer mode); the
device-resident classes at the bottom take torch CUDA tensors (device pointer mode).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked), but
creating a Context without the CUDA library and a B200 raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libqcp2_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "qcp2_b200.h")

INT_NAMES = ["oooo", "ovoo", "oovo", "ooov", "ovov", "ovvo", "oovv", "vovv", "ovvv", "vvvo", "vvvv"]
INTER_NAMES = ["F_ae", "F_mi", "F_me", "W_mnij", "W_mbej", "W_abef"]
HOST, DEVICE = 0, 1
T_AUTO, T_LITERAL, T_SYMMETRIC = 0, 1, 2
COMM_ID_BYTES = 128

_lib = None
PD = C.POINTER(C.c_double)


class QCP2Error(RuntimeError):
    pass


def load_library():
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QCP2Error(f"{LIB_PATH} is missing: build it with `make -C {HERE}` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.qcp2_version.restype = C.c_char_p
        lib.qcp2_last_error.restype = C.c_char_p
        lib.qcp2_kernel_launches.restype = C.c_longlong
        lib.qcp2_ccsd_t_count.restype = C.c_longlong
        lib.qcp2_ccsd_last_loop_seconds.restype = C.c_double
        lib.qcp2_sharded_gemms.restype = C.c_longlong
        lib.qcp2_gemm_counts.restype = C.c_longlong
        lib.qcp2_gemm_log_read.restype = C.c_longlong
        lib.qcp2_pool_high_water.restype = C.c_longlong
        _lib = lib
    return _lib


def comm_unique_id():
    """qcp2_comm_get_unique_id: the NCCL bootstrap id, created by rank 0 and handed to every rank."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    if load_library().qcp2_comm_get_unique_id(buf, COMM_ID_BYTES) != 0:
        raise QCP2Error("qcp2_comm_get_unique_id failed (is NCCL available?)")
    return buf.raw


def inter_shapes(o, v):
    return [(v, v), (o, o), (o, v), (o, o, o, o), (o, v, v, o), (v, v, v, v)]


def int_shape(spec, o, v):
    return tuple(o if c == "o" else v for c in spec)


def _arr(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(PD)
    return C.cast(C.c_void_p(a.data_ptr()), PD)  # torch tensor


def _ptrs(arrs):
    out = (PD * len(arrs))()
    for k, a in enumerate(arrs):
        out[k] = _p(a) if a is not None else None
    return out


class Context:
    """One GPU, one stream (qcp2_create / qcp2_destroy)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.qcp2_create(int(device), C.byref(self.h))
        if rc != 0:
            raise QCP2Error(f"qcp2_create(device={device}) failed with status {rc}: a CUDA B200 device is required, "
                            "there is no CPU fallback")
        self.device = device
        self._children = []  # resident plans; destroyed before the context

    def close(self):
        for child in list(getattr(self, "_children", [])):
            child.close()
        if getattr(self, "h", None) is not None and self.h:
            self.lib.qcp2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise QCP2Error(self.lib.qcp2_last_error(self.h).decode())

    def set_pointer_mode(self, mode):
        self.check(self.lib.qcp2_set_pointer_mode(self.h, mode))

    def set_stream(self, cuda_stream_ptr):
        """Enqueue device-pointer calls on the given cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream);
        0/None restores the context's own stream.  Note torch's *default* stream has handle 0: to
        share a stream with torch, run under a dedicated `torch.cuda.Stream()` and pass its handle."""
        self.check(self.lib.qcp2_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        self.check(self.lib.qcp2_synchronize(self.h))

    @property
    def kernel_launches(self):
        return int(self.lib.qcp2_kernel_launches(self.h))

    # ---- integrals.h ---------------------------------------------------------------
    def transform_ao_mo(self, pq_rs, Coeff1, Coeff2):
        ao, c1, c2 = _arr(pq_rs), _arr(Coeff1), _arr(Coeff2)
        n = ao.shape[0]
        out = np.empty((n,) * 4)
        self.check(self.lib.qcp2_transform_ao_mo(self.h, _p(ao), _p(c1), _p(c2), n, _p(out)))
        return out

    def transform_to_so(self, mo_aa, mo_bb, mo_ab):
        aa, bb, ab = _arr(mo_aa), _arr(mo_bb), _arr(mo_ab)
        n = aa.shape[0]
        out = np.empty((2 * n,) * 4)
        self.check(self.lib.qcp2_transform_to_so(self.h, _p(aa), _p(bb), _p(ab), n, _p(out)))
        return out

    def slice_ints(self, so_ints, no, nv, int_string):
        so = _arr(so_ints)
        out = np.empty(int_shape(int_string, no, nv))
        self.check(self.lib.qcp2_slice_ints(self.h, _p(so), so.shape[0], no, nv, int_string.encode(), _p(out)))
        return out

    def make_denom(self, F_spin, no, nv):
        F = _arr(F_spin)
        out = np.empty((no, no, nv, nv))
        self.check(self.lib.qcp2_make_denom(self.h, _p(F), F.shape[0], no, nv, _p(out)))
        return out

    # ---- mp2.h ---------------------------------------------------------------------
    def make_so_moes(self, eps_a, eps_b, nao=None):
        ea, eb = _arr(eps_a), _arr(eps_b)
        n = ea.shape[0] if nao is None else nao
        out = np.empty((2 * n, 2 * n))
        self.check(self.lib.qcp2_make_so_moes(self.h, _p(ea), _p(eb), n, _p(out)))
        return out

    def run_mp2(self, oovv, denom):
        g, d = _arr(oovv), _arr(denom)
        no, nv = g.shape[0], g.shape[2]
        T = np.empty_like(g)
        E = C.c_double()
        self.check(self.lib.qcp2_run_mp2(self.h, _p(g), _p(d), no, nv, _p(T), C.byref(E)))
        return T, E.value

    def MP2(self, ao_ints, Ca, Cb, eps_a, eps_b, no, nv, uhf=False, want_so=True):
        """mp2.cpp:45-86 from the AO integrals on; returns (energy, so_ints, T)."""
        ao, ca, ea = _arr(ao_ints), _arr(Ca), _arr(eps_a)
        cb, eb = (_arr(Cb), _arr(eps_b)) if uhf else (ca, ea)
        n = ao.shape[0]
        so = np.empty((2 * n,) * 4) if want_so else None
        T = np.empty((no, no, nv, nv))
        E = C.c_double()
        self.check(self.lib.qcp2_mp2(self.h, _p(ao), _p(ca), _p(cb), _p(ea), _p(eb), n, no, nv, int(bool(uhf)),
                                     _p(so), _p(T), C.byref(E)))
        return E.value, so, T

    # ---- cc.h ----------------------------------------------------------------------
    def make_fock(self, eps1, eps2, n1, n2):
        e1, e2 = _arr(eps1), _arr(eps2)
        out = np.empty((n2 - n1, n2 - n1))
        self.check(self.lib.qcp2_make_fock(self.h, _p(e1), _p(e2), n1, n2, _p(out)))
        return out

    def make_tau(self, Ts, Td, bar=False):
        ts, td = _arr(Ts), _arr(Td)
        out = np.empty_like(td)
        self.check(self.lib.qcp2_make_tau(self.h, _p(ts), _p(td), ts.shape[0], ts.shape[1], int(bar), _p(out)))
        return out

    def make_tau_bar(self, Ts, Td):
        return self.make_tau(Ts, Td, True)

    def multiply_Ts(self, Ts):
        ts = _arr(Ts)
        o, v = ts.shape
        out = np.empty((o, v, o, v))
        self.check(self.lib.qcp2_multiply_Ts(self.h, _p(ts), o, v, _p(out)))
        return out

    def _make_D(self, fn, f_oo, f_vv, shape):
        foo, fvv = _arr(f_oo), _arr(f_vv)
        out = np.empty(shape)
        self.check(fn(self.h, _p(foo), _p(fvv), foo.shape[0], fvv.shape[0], _p(out)))
        return out

    def make_D_ia(self, f_oo, f_vv):
        return self._make_D(self.lib.qcp2_make_D_ia, f_oo, f_vv, (len(f_oo), len(f_vv)))

    def make_D_ijab(self, f_oo, f_vv):
        o, v = len(f_oo), len(f_vv)
        return self._make_D(self.lib.qcp2_make_D_ijab, f_oo, f_vv, (o, o, v, v))

    def make_D_triples(self, f_oo, f_vv):
        o, v = len(f_oo), len(f_vv)
        return self._make_D(self.lib.qcp2_make_D_triples, f_oo, f_vv, (o, o, o, v, v, v))

    def update_intermediates(self, Ts, Td, integrals, f_oo, f_ov, f_vv):
        ts, td, foo, fov, fvv = _arr(Ts), _arr(Td), _arr(f_oo), _arr(f_ov), _arr(f_vv)
        o, v = ts.shape
        ints = [_arr(integrals.get(k)) for k in INT_NAMES]
        outs = [np.empty(s) for s in inter_shapes(o, v)]
        self.check(self.lib.qcp2_update_intermediates(self.h, _p(ts), _p(td), _ptrs(ints), _p(foo), _p(fov), _p(fvv),
                                                      o, v, _ptrs(outs)))
        return dict(zip(INTER_NAMES, outs))

    def make_T1(self, Ts, Td, integrals, intermediates, f_oo, f_ov, f_vv, D_ia=None):
        ts, td, foo, fov, fvv, dia = _arr(Ts), _arr(Td), _arr(f_oo), _arr(f_ov), _arr(f_vv), _arr(D_ia)
        o, v = ts.shape
        ints = [_arr(integrals.get(k)) for k in INT_NAMES]
        inter = [_arr(intermediates.get(k)) for k in INTER_NAMES]
        out = np.empty((o, v))
        self.check(self.lib.qcp2_make_T1(self.h, _p(ts), _p(td), _ptrs(ints), _ptrs(inter), _p(foo), _p(fov), _p(fvv),
                                         _p(dia), o, v, _p(out)))
        return out

    def make_T2(self, Ts, Td, integrals, intermediates, f_oo=None, f_vv=None, D_ijab=None):
        ts, td, foo, fvv, dd = _arr(Ts), _arr(Td), _arr(f_oo), _arr(f_vv), _arr(D_ijab)
        o, v = ts.shape
        ints = [_arr(integrals.get(k)) for k in INT_NAMES]
        inter = [_arr(intermediates.get(k)) for k in INTER_NAMES]
        out = np.empty((o, o, v, v))
        self.check(self.lib.qcp2_make_T2(self.h, _p(ts), _p(td), _ptrs(ints), _ptrs(inter), _p(foo), _p(fvv), _p(dd),
                                         o, v, _p(out)))
        return out

    def ccsd_energy(self, ts, td, oovv, f_ov):
        ts, td, g, fov = _arr(ts), _arr(td), _arr(oovv), _arr(f_ov)
        E = C.c_double()
        self.check(self.lib.qcp2_ccsd_energy(self.h, _p(ts), _p(td), _p(g), _p(fov), ts.shape[0], ts.shape[1],
                                             C.byref(E)))
        return E.value

    def CCSD(self, f_oo, f_ov, f_vv, T2_guess, maxiter, cc_conv, so_ints=None, integrals=None, want_sliced=False):
        """cc.cpp:409-451.  Returns dict(T1, T2, ccsd_energy, iters, e_hist, loop_seconds[, sliced_ints])."""
        foo, fov, fvv, guess = _arr(f_oo), _arr(f_ov), _arr(f_vv), _arr(T2_guess)
        o, v = foo.shape[0], fvv.shape[0]
        so = _arr(so_ints)
        ints = [_arr(integrals.get(k)) for k in INT_NAMES] if integrals is not None else None
        T1, T2 = np.empty((o, v)), np.empty((o, o, v, v))
        E, iters = C.c_double(), C.c_int()
        hist = np.zeros(max(maxiter, 1))
        sliced = [np.empty(int_shape(k, o, v)) for k in INT_NAMES] if want_sliced else None
        self.check(self.lib.qcp2_ccsd(self.h, _p(so), 0 if so is None else so.shape[0],
                                      _ptrs(ints) if ints is not None else None, _p(foo), _p(fov), _p(fvv), _p(guess),
                                      o, v, maxiter, C.c_double(cc_conv), _p(T1), _p(T2), C.byref(E), C.byref(iters),
                                      _p(hist), _ptrs(sliced) if sliced is not None else None))
        res = dict(T1=T1, T2=T2, ccsd_energy=E.value, iters=iters.value,
                   e_hist=hist[:min(iters.value + 1, maxiter)],
                   loop_seconds=float(self.lib.qcp2_ccsd_last_loop_seconds(self.h)))
        if want_sliced:
            res["sliced_ints"] = dict(zip(INT_NAMES, sliced))
        return res

    def post_hf(self, ao_ints, Ca, Cb, eps_a, eps_b, no, nv, uhf=False, stages=3, maxiter=200, cc_conv=1e-12,
                want_amplitudes=False):
        """Fused device-resident MP2 -> CCSD -> (T) (qcp2_post_hf): what main.cpp:76-88 runs after the SCF."""
        ao, ca, ea = _arr(ao_ints), _arr(Ca), _arr(eps_a)
        cb, eb = (_arr(Cb), _arr(eps_b)) if uhf else (None, None)
        n = ao.shape[0]
        e2, ecc, et, iters = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        hist = np.zeros(max(maxiter, 1))
        T1 = np.empty((no, nv)) if want_amplitudes else None
        T2 = np.empty((no, no, nv, nv)) if want_amplitudes else None
        self.check(self.lib.qcp2_post_hf(self.h, _p(ao), _p(ca), _p(cb), _p(ea), _p(eb), n, no, nv, int(bool(uhf)), stages,
                                         maxiter, C.c_double(cc_conv), C.byref(e2), C.byref(ecc), C.byref(et),
                                         C.byref(iters), _p(hist), _p(T1), _p(T2)))
        out = dict(e_mp2=e2.value, e_ccsd=ecc.value, e_t=et.value, iters=iters.value,
                   e_hist=hist[:min(iters.value + 1, maxiter)])
        if want_amplitudes:
            out.update(T1=T1, T2=T2)
        return out

    # ---- multi-GPU: NCCL communicator owned by the library (one rank per GPU) ----------------
    def comm_init(self, unique_id, nranks, rank):
        """qcp2_comm_init: collective; unique_id = the 128 bytes rank 0 got from comm_unique_id()."""
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self.check(self.lib.qcp2_comm_init(self.h, buf, int(nranks), int(rank)))

    def comm_init_from_torch(self, dist):
        """Bootstrap the library's own NCCL communicator over an initialised torch.distributed group:
        rank 0 creates the unique id, the host channel of the group hands it round."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(box[0], world, rank)

    def comm_destroy(self):
        self.check(self.lib.qcp2_comm_destroy(self.h))

    def comm_rank(self):
        r, n = C.c_int(), C.c_int()
        self.lib.qcp2_comm_rank(self.h, C.byref(r), C.byref(n))
        return r.value, n.value

    def comm_broadcast(self, dev_tensor, root=0):
        self.check(self.lib.qcp2_comm_broadcast(self.h, _p(dev_tensor), C.c_longlong(dev_tensor.numel()), int(root)))

    def comm_allreduce_sum(self, dev_tensor):
        self.check(self.lib.qcp2_comm_allreduce_sum(self.h, _p(dev_tensor), C.c_longlong(dev_tensor.numel())))

    def comm_allgather(self, dev_send, dev_recv):
        self.check(self.lib.qcp2_comm_allgather(self.h, _p(dev_send), _p(dev_recv), C.c_longlong(dev_send.numel())))

    def CCSD_T_mgpu(self, T1, T2, oovv, vovv, ovoo, eps_o, eps_v, o, v, mode=T_AUTO, root=0, want_triples=False):
        """cc.cpp:454-560 across the communicator (qcp2_ccsd_t_mgpu): the root rank passes the tensors
        (numpy in host pointer mode, torch CUDA tensors in device pointer mode), the others None."""
        conv = (lambda a: a) if (T1 is not None and not isinstance(T1, np.ndarray)) else _arr
        args = [conv(a) if a is not None else None for a in (T1, T2, oovv, vovv, ovoo, eps_o, eps_v)]
        E = C.c_double()
        et = np.zeros(max(self.t_count(o, T_SYMMETRIC if mode != T_LITERAL else T_LITERAL), 1)) if want_triples else None
        if want_triples and mode == T_AUTO:
            et = np.zeros(max(self.t_count(o, T_LITERAL), 1))  # large enough for either enumeration
        self.check(self.lib.qcp2_ccsd_t_mgpu(self.h, *[_p(a) for a in args], int(o), int(v), int(mode), int(root),
                                             _p(et), C.byref(E)))
        return (E.value, et) if want_triples else E.value

    def post_hf_mgpu(self, ao_ints, Ca, Cb, eps_a, eps_b, n, no, nv, uhf=False, stages=3, maxiter=200, cc_conv=1e-12, root=0):
        """qcp2_post_hf_mgpu (collective): the root rank passes the boundary inputs, the others None."""
        conv = (lambda a: a) if (ao_ints is not None and not isinstance(ao_ints, np.ndarray)) else _arr
        ao, ca, ea = (conv(x) if x is not None else None for x in (ao_ints, Ca, eps_a))
        cb, eb = ((conv(Cb), conv(eps_b)) if (uhf and Cb is not None) else (None, None))
        e2, ecc, et, iters = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        hist = np.zeros(max(maxiter, 1))
        self.check(self.lib.qcp2_post_hf_mgpu(self.h, _p(ao), _p(ca), _p(cb), _p(ea), _p(eb), int(n), int(no), int(nv), int(bool(uhf)),
                                              int(stages), int(maxiter), C.c_double(cc_conv), int(root), C.byref(e2), C.byref(ecc),
                                              C.byref(et), C.byref(iters), _p(hist), None, None))
        return dict(e_mp2=e2.value, e_ccsd=ecc.value, e_t=et.value, iters=iters.value, e_hist=hist[:min(iters.value + 1, maxiter)])

    def set_ccsd_sharding(self, enable):
        self.check(self.lib.qcp2_set_ccsd_sharding(self.h, int(bool(enable))))

    @property
    def sharded_gemms(self):
        return int(self.lib.qcp2_sharded_gemms(self.h))

    def t_count(self, o, mode):
        return int(self.lib.qcp2_ccsd_t_count(o, mode))

    def t_select_mode(self, T2, oovv, vovv, ovoo):
        t2, g, w, x = _arr(T2), _arr(oovv), _arr(vovv), _arr(ovoo)
        o, v = t2.shape[0], t2.shape[2]
        m = C.c_int()
        self.check(self.lib.qcp2_ccsd_t_select_mode(self.h, _p(t2), _p(g), _p(w), _p(x), o, v, C.byref(m)))
        return m.value

    def CCSD_T(self, T1, T2, oovv, vovv, ovoo, eps_o, eps_v, mode=T_AUTO):
        """cc.cpp:454-560."""
        t1, t2, g, w, x, eo, ev = map(_arr, (T1, T2, oovv, vovv, ovoo, eps_o, eps_v))
        E = C.c_double()
        self.check(self.lib.qcp2_ccsd_t(self.h, _p(t1), _p(t2), _p(g), _p(w), _p(x), _p(eo), _p(ev), t1.shape[0],
                                        t1.shape[1], mode, C.byref(E)))
        return E.value

    def CCSD_T_range(self, T1, T2, oovv, vovv, ovoo, eps_o, eps_v, mode, first, stride, max_count=-1, e_triples=None):
        t1, t2, g, w, x, eo, ev = map(_arr, (T1, T2, oovv, vovv, ovoo, eps_o, eps_v))
        E = C.c_double()
        self.check(self.lib.qcp2_ccsd_t_range(self.h, _p(t1), _p(t2), _p(g), _p(w), _p(x), _p(eo), _p(ev), t1.shape[0],
                                              t1.shape[1], mode, C.c_longlong(first), C.c_longlong(stride),
                                              C.c_longlong(max_count), _p(e_triples), C.byref(E)))
        return E.value

    # ---- building blocks -----------------------------------------------------------
    def contract(self, alpha, A, ia, B, ib, beta, Cin, ic):
        a, b, c = _arr(A), _arr(B), _arr(Cin).copy()
        ext = lambda t: (C.c_longlong * t.ndim)(*t.shape)
        self.check(self.lib.qcp2_contract(self.h, C.c_double(alpha), _p(a), ext(a), ia.encode(), _p(b), ext(b),
                                          ib.encode(), C.c_double(beta), _p(c), ext(c), ic.encode()))
        return c

    def dgemm_device(self, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, Cdev, ldc):
        """Raw device-pointer GEMM (torch CUDA tensors); pointer mode must be DEVICE."""
        self.check(self.lib.qcp2_dgemm(self.h, int(transA), int(transB), M, N, K, C.c_double(alpha), _p(A),
                                       C.c_longlong(lda), _p(B), C.c_longlong(ldb), C.c_double(beta), _p(Cdev),
                                       C.c_longlong(ldc)))

    def dgemm_ex(self, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, Cdev, ldc, batch=1, strideA=0, strideB=0,
                 strideC=0, tile=-1, k_splits=0, reps=1):
        """qcp2_dgemm_ex: device pointers, strided batch, forced tile / split-K of the warp-specialised TMA kernel;
        returns the mean device seconds of one launch."""
        secs = C.c_double()
        self.check(self.lib.qcp2_dgemm_ex(self.h, int(transA), int(transB), M, N, K, C.c_double(alpha), _p(A), C.c_longlong(lda),
                                          _p(B), C.c_longlong(ldb), C.c_double(beta), _p(Cdev), C.c_longlong(ldc), int(batch),
                                          C.c_longlong(strideA), C.c_longlong(strideB), C.c_longlong(strideC), int(tile),
                                          int(k_splits), int(reps), C.byref(secs)))
        return secs.value

    def so_block(self, mo_aa, mo_bb, mo_ab, no, nv, int_string):
        """qcp2_so_block: slice_ints(transform_to_so(..)) gathered straight from the spatial MO tensors."""
        aa, bb, ab = _arr(mo_aa), _arr(mo_bb), _arr(mo_ab)
        out = np.empty(int_shape(int_string, no, nv))
        self.check(self.lib.qcp2_so_block(self.h, _p(aa), _p(bb), _p(ab), aa.shape[0], no, nv, int_string.encode(), _p(out)))
        return out

    def ccsd_set_start_T1(self, T1):
        """The next CCSD solve starts from this T1 instead of zero (restart from saved amplitudes); None cancels."""
        if T1 is None:
            self.check(self.lib.qcp2_ccsd_set_start_T1(self.h, None, 0, 0))
        else:
            t = _arr(T1) if isinstance(T1, np.ndarray) else T1
            self.check(self.lib.qcp2_ccsd_set_start_T1(self.h, _p(t), int(t.shape[0]), int(t.shape[1])))

    def set_ccsd_diis(self, nvec):
        """Opt-in DIIS over the last nvec Jacobi steps (0 = the reference's plain Jacobi loop)."""
        self.check(self.lib.qcp2_set_ccsd_diis(self.h, int(nvec)))

    def ccsd_last_mode(self):
        """0 dense term list, 1 pair-packed, 2 pair-packed + spin classes (see qcp2_ccsd_last_mode)."""
        return int(self.lib.qcp2_ccsd_last_mode(self.h))

    def pool_high_water(self):
        """Peak bytes of the context's workspace pool since the start of the last post_hf call."""
        return int(self.lib.qcp2_pool_high_water(self.h))

    def gemm_log(self, enable=True):
        self.check(self.lib.qcp2_gemm_log(self.h, int(bool(enable))))

    def gemm_log_read(self):
        """Products of the last recorded CCSD iteration: list of dicts M, N, K, batch, beta, tile_m, tile_n, split."""
        n = int(self.lib.qcp2_gemm_log_read(self.h, None, C.c_longlong(0)))
        buf = (C.c_longlong * (8 * max(n, 1)))()
        self.lib.qcp2_gemm_log_read(self.h, buf, C.c_longlong(n))
        keys = ("M", "N", "K", "batch", "beta", "tile_m", "tile_n", "split")
        return [dict(zip(keys, buf[8 * k:8 * k + 8])) for k in range(n)]

    def gemm_counts(self):
        """(products on the warp-specialised TMA kernel, products on the cp.async kernel) so far."""
        legacy = C.c_longlong()
        ws = int(self.lib.qcp2_gemm_counts(self.h, C.byref(legacy)))
        return ws, int(legacy.value)

    def dgemm(self, transA, transB, alpha, A, B, beta, Cin):
        a, b, c = _arr(A), _arr(B), _arr(Cin).copy()
        M, N = c.shape
        K = a.shape[0] if transA else a.shape[1]
        self.check(self.lib.qcp2_dgemm(self.h, int(transA), int(transB), M, N, K, C.c_double(alpha), _p(a),
                                       C.c_longlong(a.shape[1]), _p(b), C.c_longlong(b.shape[1]), C.c_double(beta),
                                       _p(c), C.c_longlong(c.shape[1])))
        return c


# --------------------------------------------------------------------------------------
# device-resident helpers (torch CUDA tensors; used by bench.py and the multi-GPU driver)
# --------------------------------------------------------------------------------------
class DeviceTriples:
    """Resident (T) plan: inputs stay in HBM, range evaluations reuse the packed tensors."""

    def __init__(self, ctx, T1, T2, oovv, vovv, ovoo, eps_o, eps_v, mode=T_SYMMETRIC):
        self.ctx = ctx
        self.keep = (T1, T2, oovv, vovv, ovoo, eps_o, eps_v)
        o, v = T1.shape
        self.o, self.v = o, v
        ctx.set_pointer_mode(DEVICE)
        self.h = C.c_void_p()
        if hasattr(T1, "is_cuda"):
            import torch
            torch.cuda.current_stream().synchronize()  # inputs may have been produced on a torch stream
        ctx.check(ctx.lib.qcp2_t_plan_create(ctx.h, _p(T1), _p(T2), _p(oovv), _p(vovv), _p(ovoo), _p(eps_o),
                                             _p(eps_v), o, v, mode, C.byref(self.h)))
        ctx._children.append(self)

    def run(self, first=0, stride=1, max_count=-1, e_triples=None):
        """e_triples: torch CUDA tensor of t_count() doubles (or None).  The caller's torch stream is
        synchronised first so tensors prepared there (zero_(), broadcast) are complete even when the
        context runs on its own stream; the call itself returns after the device work is done."""
        if e_triples is not None and hasattr(e_triples, "is_cuda"):
            import torch
            torch.cuda.current_stream().synchronize()
        E = C.c_double()
        self.ctx.check(self.ctx.lib.qcp2_t_plan_run(self.h, C.c_longlong(first), C.c_longlong(stride),
                                                    C.c_longlong(max_count), _p(e_triples), C.byref(E)))
        return E.value

    def info(self):
        tma, tb = C.c_int(), C.c_int()
        nc, ld = C.c_longlong(), C.c_longlong()
        self.ctx.check(self.ctx.lib.qcp2_t_plan_info(self.h, C.byref(tma), C.byref(nc), C.byref(ld), C.byref(tb)))
        return {"uses_tma_kernel": bool(tma.value), "ncols": nc.value, "row_stride": ld.value, "triples_per_batch": tb.value}

    def timings(self):
        g, t = C.c_double(), C.c_double()
        ng, na = C.c_longlong(), C.c_longlong()
        self.ctx.lib.qcp2_t_plan_timings(self.h, C.byref(g), C.byref(t), C.byref(ng), C.byref(na))
        return dict(gemm_seconds=g.value, total_seconds=t.value, gemm_launches=ng.value, all_launches=na.value)

    def close(self):
        if self.h:
            self.ctx.lib.qcp2_t_plan_destroy(self.h)
            self.h = None
            if self in self.ctx._children:
                self.ctx._children.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_t_inputs_device(ctx, o, v, seed=20261019, device="cuda"):
    """SURVEY.md section 8d synthetic (T) inputs generated directly in HBM (torch tensors)."""
    import torch
    f = dict(dtype=torch.float64, device=device)
    T1, T2 = torch.empty((o, v), **f), torch.empty((o, o, v, v), **f)
    oovv, vovv, ovoo = torch.empty((o, o, v, v), **f), torch.empty((v, o, v, v), **f), torch.empty((o, v, o, o), **f)
    eo, ev = torch.empty(o, **f), torch.empty(v, **f)
    torch.cuda.synchronize()
    ctx.check(ctx.lib.qcp2_synth_t_inputs(ctx.h, o, v, C.c_ulonglong(seed), _p(T1), _p(T2), _p(oovv), _p(vovv),
                                          _p(ovoo), _p(eo), _p(ev)))
    return dict(T1=T1, T2=T2, oovv=oovv, vovv=vovv, ovoo=ovoo, eps_o=eo, eps_v=ev)
