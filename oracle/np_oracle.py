"""oracle/np_oracle.py -- TEST INFRASTRUCTURE ONLY (second, independent CPU oracle + CPU baseline).

numpy/einsum restatement of the post-HF hot path of ajay-mk/QC-P2, written from the
equations the reference cites (Stanton, Gauss, Watts, Bartlett, JCP 94, 4334 (1991)
eqs. 1-13; Crawford ProgrammingProjects #6 for (T)) rather than from oracle.cpp, so the two
oracles check each other.  Reference call sites: src/integrals.cpp:95-207,
src/mp2.cpp:11-42, src/cc.cpp:125-560.

Pinned through oracle.cpp, which tests/test_ref_pins_oracle.py holds against the reference's own
sources compiled under oracle/_ref (see oracle/ref.py), and to the external Crawford
water/STO-3G numbers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import itertools

import functools

import numpy as np

MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_es = functools.partial(np.einsum, optimize=True)  # contractions go through tensordot/BLAS


# --------------------------------------------------------------------------------------
# counter-based synthetic data (same generator as csrc/synth.cu)
# --------------------------------------------------------------------------------------
def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def u_rand(seed, tag, idx):
    """u(tag, idx) in [-1, 1): (splitmix64(seed ^ tag<<56 ^ idx) >> 11) * 2^-52 - 1."""
    key = np.uint64(seed) ^ (np.uint64(tag) << np.uint64(56))
    z = splitmix64(np.asarray(idx, dtype=np.uint64) ^ key)
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0


TAG_T1, TAG_T2, TAG_OOVV, TAG_VOVV, TAG_OVOO = 1, 2, 3, 4, 5
SEED = 20261019


def _sorted_pair(x, y):
    """(lo, hi, sign) of an index pair: sign +1 if x<y, -1 if x>y, 0 on the diagonal."""
    return np.minimum(x, y), np.maximum(x, y), np.sign(y - x).astype(np.float64)


def _oovv_like_rows(o, v, tag, scale, seed, rows):
    """Rows `rows` of an [o,o,v,v] tensor antisymmetric in (0,1) and (2,3): value =
    sign * scale * u(tag, linear index of the canonical i<j, a<b element)."""
    out = np.zeros((o, o, v, v))
    j = np.arange(o, dtype=np.int64).reshape(-1, 1, 1)
    a = np.arange(v, dtype=np.int64).reshape(1, -1, 1)
    b = np.arange(v, dtype=np.int64).reshape(1, 1, -1)
    alo, ahi, asg = _sorted_pair(a, b)
    for i in rows:
        ilo, ihi, isg = _sorted_pair(np.int64(i), j)
        idx = ((ilo * o + ihi) * v + alo) * v + ahi
        out[i] = (isg * asg) * (scale * u_rand(seed, tag, idx.astype(np.uint64)))
    return out


def synthetic_t_inputs(o, v, seed=SEED, occ=None):
    """Antisymmetrised random (T) inputs of SURVEY.md section 8d (config 5 generator).
    occ = iterable of occupied indices: only the slabs a triple over those indices reads are
    generated (T2[occ], oovv[occ], vovv[:, occ]); the rest of the (lazily allocated, zero)
    full-shape arrays stays untouched, so full-size samples cost a few GB instead of 25."""
    rows = list(range(o)) if occ is None else sorted(set(int(x) for x in occ))
    T1 = 0.05 * u_rand(seed, TAG_T1, np.arange(o * v, dtype=np.uint64)).reshape(o, v)
    T2 = _oovv_like_rows(o, v, TAG_T2, 0.02, seed, rows)
    oovv = _oovv_like_rows(o, v, TAG_OOVV, 0.1, seed, rows)
    vovv = np.zeros((v, o, v, v))
    e = np.arange(v, dtype=np.int64).reshape(-1, 1, 1)
    b = np.arange(v, dtype=np.int64).reshape(1, -1, 1)
    c = np.arange(v, dtype=np.int64).reshape(1, 1, -1)
    blo, bhi, bsg = _sorted_pair(b, c)
    for i in rows:
        idx = ((e * o + i) * v + blo) * v + bhi
        vovv[:, i] = bsg * (0.1 * u_rand(seed, TAG_VOVV, idx.astype(np.uint64)))
    m = np.arange(o, dtype=np.int64).reshape(-1, 1, 1, 1)
    a = np.arange(v, dtype=np.int64).reshape(1, -1, 1, 1)
    jj = np.arange(o, dtype=np.int64).reshape(1, 1, -1, 1)
    kk = np.arange(o, dtype=np.int64).reshape(1, 1, 1, -1)
    jlo, jhi, jsg = _sorted_pair(jj, kk)
    ovoo = jsg * (0.1 * u_rand(seed, TAG_OVOO, (((m * v + a) * o + jlo) * o + jhi).astype(np.uint64)))
    eo = -2.0 + 1.5 * np.arange(o) / max(o - 1, 1)
    ev = 0.1 + 2.9 * np.arange(v) / max(v - 1, 1)
    return dict(T1=T1, T2=T2, oovv=oovv, vovv=vovv, ovoo=ovoo, eps_o=eo, eps_v=ev)


# --------------------------------------------------------------------------------------
# integrals.cpp / mp2.cpp
# --------------------------------------------------------------------------------------
def transform_ao_mo(ao, C1, C2):
    """(pq|rs) -> (ij|kl): k,l rotated by C1, i,j by C2 (integrals.cpp:121-129)."""
    return np.einsum("pqrs,pi,qj,rk,sl->ijkl", ao, C2, C2, C1, C1, optimize=True)


def transform_to_so(aa, bb, ab):
    """integrals.cpp:136-167, including the as-written alpha/beta label use."""
    n = aa.shape[0]
    so = np.zeros((2 * n,) * 4)
    A, B = slice(0, None, 2), slice(1, None, 2)
    phys = lambda t: t.transpose(0, 2, 1, 3)  # t[I,K,J,L] -> [I,J,K,L]
    so[A, A, A, A] = phys(aa) - phys(aa).transpose(1, 0, 2, 3)
    so[B, B, B, B] = phys(bb) - phys(bb).transpose(1, 0, 2, 3)
    so[A, B, A, B] = phys(ab)                      # i a, k a, j b, l b
    so[B, A, B, A] = phys(ab)                      # i b, k b, j a, l a
    so[B, A, A, B] = -phys(ab).transpose(1, 0, 2, 3)  # -(ab)[J,K,I,L]
    so[A, B, B, A] = -phys(ab).transpose(1, 0, 2, 3)
    return so


def slice_ints(so, no, spec):
    sl = [slice(0, no) if c == "o" else slice(no, None) for c in spec]
    return np.ascontiguousarray(so[tuple(sl)])


INT_NAMES = ["oooo", "ovoo", "oovo", "ooov", "ovov", "ovvo", "oovv", "vovv", "ovvv", "vvvo", "vvvv"]


def get_integrals(so, no):
    return {k: slice_ints(so, no, k) for k in INT_NAMES}


def mp2(oovv, eo, ev):
    D = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
    return oovv / D, 0.25 * np.sum(oovv * oovv / D)


# --------------------------------------------------------------------------------------
# cc.cpp: CCSD (Stanton 1991), written with physicists' einsum
# --------------------------------------------------------------------------------------
def _taus(t1, t2):
    tt = _es("ia,jb->ijab", t1, t1)
    tt = tt - tt.transpose(0, 1, 3, 2)
    return t2 + tt, t2 + 0.5 * tt


def intermediates(t1, t2, I, foo, fov, fvv):
    tau, taub = _taus(t1, t2)
    offd = lambda f: f - np.diag(np.diag(f))
    Fae = offd(fvv) - 0.5 * _es("me,ma->ae", fov, t1) + _es("mf,mafe->ae", t1, I["ovvv"]) \
        - 0.5 * _es("mnaf,mnef->ae", taub, I["oovv"])
    Fmi = offd(foo) + 0.5 * _es("ie,me->mi", t1, fov) + _es("ne,mnie->mi", t1, I["ooov"]) \
        + 0.5 * _es("inef,mnef->mi", taub, I["oovv"])
    Fme = fov + _es("nf,mnef->me", t1, I["oovv"])
    Wmnij = I["oooo"] + _es("je,mnie->mnij", t1, I["ooov"]) - _es("ie,mnje->mnij", t1, I["ooov"]) \
        + 0.25 * _es("ijef,mnef->mnij", tau, I["oovv"])
    Wabef = I["vvvv"] - _es("mb,amef->abef", t1, I["vovv"]) + _es("ma,bmef->abef", t1, I["vovv"]) \
        + 0.25 * _es("mnab,mnef->abef", tau, I["oovv"])
    Wmbej = I["ovvo"] + _es("jf,mbef->mbej", t1, I["ovvv"]) - _es("nb,mnej->mbej", t1, I["oovo"]) \
        - _es("jnfb,mnef->mbej", 0.5 * t2 + _es("jf,nb->jnfb", t1, t1), I["oovv"])
    return dict(F_ae=Fae, F_mi=Fmi, F_me=Fme, W_mnij=Wmnij, W_mbej=Wmbej, W_abef=Wabef)


def new_T1(t1, t2, I, X, fov, eo, ev):
    r = fov + _es("ie,ae->ia", t1, X["F_ae"]) - _es("ma,mi->ia", t1, X["F_mi"]) \
        + _es("imae,me->ia", t2, X["F_me"]) - _es("nf,naif->ia", t1, I["ovov"]) \
        - 0.5 * _es("imef,maef->ia", t2, I["ovvv"]) - 0.5 * _es("mnae,nmei->ia", t2, I["oovo"])
    return r / (eo[:, None] - ev[None, :])


def new_T2(t1, t2, I, X, eo, ev):
    tau, _ = _taus(t1, t2)
    Pab = lambda x: x - x.transpose(0, 1, 3, 2)
    Pij = lambda x: x - x.transpose(1, 0, 2, 3)
    r = I["oovv"].copy()
    r += Pab(_es("ijae,be->ijab", t2, X["F_ae"] - 0.5 * _es("mb,me->be", t1, X["F_me"])))
    r -= Pij(_es("imab,mj->ijab", t2, X["F_mi"] + 0.5 * _es("je,me->mj", t1, X["F_me"])))
    r += 0.5 * _es("mnab,mnij->ijab", tau, X["W_mnij"])
    r += 0.5 * _es("ijef,abef->ijab", tau, X["W_abef"])
    r += Pij(Pab(_es("imae,mbej->ijab", t2, X["W_mbej"])
                 - _es("ie,ma,mbej->ijab", t1, t1, I["ovvo"])))
    r += Pij(_es("ie,abej->ijab", t1, I["vvvo"]))
    r -= Pab(_es("ma,mbij->ijab", t1, I["ovoo"]))
    D = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
    return r / D


def ccsd_energy(t1, t2, oovv, fov):
    return np.sum(fov * t1) + 0.25 * np.sum(oovv * t2) + 0.5 * _es("ijab,ia,jb->", oovv, t1, t1)


def ccsd(I, foo, fov, fvv, t2_guess, maxiter, conv):
    """cc.cpp:409-451 trajectory: E before update; new T1 feeds new_T2 with old intermediates."""
    eo, ev = np.diag(foo).copy(), np.diag(fvv).copy()
    t1 = np.zeros(fov.shape)
    t2 = t2_guess.copy()
    e_last, hist = 0.0, []
    for it in range(maxiter):
        e = ccsd_energy(t1, t2, I["oovv"], fov)
        hist.append(e)
        if abs(e - e_last) < conv:
            break
        e_last = e
        X = intermediates(t1, t2, I, foo, fov, fvv)
        t1 = new_T1(t1, t2, I, X, fov, eo, ev)
        t2 = new_T2(t1, t2, I, X, eo, ev)
    return t1, t2, hist


# note on new_T2's "- t_ie t_ma <mb||ej>" term: cc.cpp:360-363 contracts (t_ie t_ma) with
# ovov{m,b,j,e}; <mb||je> = -<mb||ej>, so +ovov there equals -ovvo here.  Only valid when
# the integrals are antisymmetric in the ket (true for RHF inputs, not for the as-written UHF
# path), so tests that use this module for UHF-as-written compare against oracle.cpp instead.


# --------------------------------------------------------------------------------------
# cc.cpp:454-560: (T)
# --------------------------------------------------------------------------------------
def ccsd_t_einsum(T1, T2, oovv, vovv, ovoo, eo, ev):
    """Crawford project #6 formulas, dense o^3 v^3 (small sizes only)."""
    o, v = T1.shape

    def perm_abc(x):  # P(a/bc): x - (a<->b) - (a<->c)
        return x - x.transpose(0, 1, 2, 4, 3, 5) - x.transpose(0, 1, 2, 5, 4, 3)

    def perm_ijk(x):  # P(i/jk)
        return x - x.transpose(1, 0, 2, 3, 4, 5) - x.transpose(2, 1, 0, 3, 4, 5)

    Dd = perm_ijk(perm_abc(np.einsum("ia,jkbc->ijkabc", T1, oovv)))
    part = np.einsum("jkae,eibc->ijkabc", T2, vovv) - np.einsum("imbc,majk->ijkabc", T2, ovoo)
    Dc = perm_ijk(perm_abc(part))
    D = (eo[:, None, None, None, None, None] + eo[None, :, None, None, None, None]
         + eo[None, None, :, None, None, None] - ev[None, None, None, :, None, None]
         - ev[None, None, None, None, :, None] - ev[None, None, None, None, None, :])
    return np.sum(Dc * (Dc + Dd) / D) / 36.0


def sorted_triples(o):
    return np.array(list(itertools.combinations(range(o), 3)), dtype=np.int32).reshape(-1, 3)


def all_triples(o):
    return np.array(list(itertools.product(range(o), repeat=3)), dtype=np.int32).reshape(-1, 3)


def ccsd_t_triple_np(T1, T2, oovv, vovv, ovoo, eo, ev, i, j, k):
    """One occupied triple via the single-GEMM restatement:
    X = A.B, K = 3(v+o); W = X - X[b,a,c] - X[c,b,a]; returns sum_abc W (W+V)/D (no 1/36)."""
    o, v = T1.shape
    A = np.concatenate([T2[j, k], -T2[i, k], -T2[j, i],
                        -ovoo[:, :, j, k].T, ovoo[:, :, i, k].T, ovoo[:, :, j, i].T], axis=1)
    B = np.concatenate([vovv[:, i].reshape(v, v * v), vovv[:, j].reshape(v, v * v), vovv[:, k].reshape(v, v * v),
                        T2[i].reshape(o, v * v), T2[j].reshape(o, v * v), T2[k].reshape(o, v * v)], axis=0)
    X = (A @ B).reshape(v, v, v)
    W = X - X.transpose(1, 0, 2) - X.transpose(2, 1, 0)
    Y = (np.einsum("a,bc->abc", T1[i], oovv[j, k]) - np.einsum("a,bc->abc", T1[j], oovv[i, k])
         - np.einsum("a,bc->abc", T1[k], oovv[j, i]))
    V = Y - Y.transpose(1, 0, 2) - Y.transpose(2, 1, 0)
    D = eo[i] + eo[j] + eo[k] - ev[:, None, None] - ev[None, :, None] - ev[None, None, :]
    return float(np.sum(W * (W + V) / D))


def ccsd_t_blocked(T1, T2, oovv, vovv, ovoo, eo, ev, symmetric):
    """Full E(T) triple by triple.  symmetric=True: i<j<k only, weight 6 (needs antisymmetric
    inputs); False: all o^3 ordered triples (literal sum)."""
    o = T1.shape[0]
    trip = sorted_triples(o) if symmetric else all_triples(o)
    w = 6.0 if symmetric else 1.0
    e = 0.0
    for (i, j, k) in trip:
        e += w * ccsd_t_triple_np(T1, T2, oovv, vovv, ovoo, eo, ev, int(i), int(j), int(k))
    return e / 36.0


def ccsd_t_triple_packed_np(T1, T2, oovv, vovv, ovoo, eo, ev, i, j, k, bu=None, cu=None, fast=False):
    """CPU baseline kernel for config 5: one i<j<k triple with b<c column packing
    (M=v, N=v(v-1)/2, K=3(v+o) GEMM via BLAS) and the a<b<c energy sum; returns the triple's
    contribution to E(T) (already including the 6*6/36 weight).  fast=True hands the energy sum to
    the OpenMP loop in oracle.cpp (all host threads) instead of whole-cube numpy passes."""
    o, v = T1.shape
    if bu is None:
        bu, cu = np.triu_indices(v, 1)
    A = np.concatenate([T2[j, k], -T2[i, k], -T2[j, i],
                        -ovoo[:, :, j, k].T, ovoo[:, :, i, k].T, ovoo[:, :, j, i].T], axis=1)
    B = np.concatenate([vovv[:, i][:, bu, cu], vovv[:, j][:, bu, cu], vovv[:, k][:, bu, cu],
                        T2[i][:, bu, cu], T2[j][:, bu, cu], T2[k][:, bu, cu]], axis=0)
    Xp = A @ B  # [v, NP]
    if fast:
        from oracle import oracle as _orc
        return _orc.t_energy_packed(Xp, T1, oovv, eo, ev, i, j, k)
    X = np.zeros((v, v, v))
    X[:, bu, cu] = Xp
    X[:, cu, bu] = -Xp
    W = X - X.transpose(1, 0, 2) - X.transpose(2, 1, 0)
    Y = (np.einsum("a,bc->abc", T1[i], oovv[j, k]) - np.einsum("a,bc->abc", T1[j], oovv[i, k])
         - np.einsum("a,bc->abc", T1[k], oovv[j, i]))
    V = Y - Y.transpose(1, 0, 2) - Y.transpose(2, 1, 0)
    D = eo[i] + eo[j] + eo[k] - ev[:, None, None] - ev[None, :, None] - ev[None, None, :]
    # W(W+V)/D is totally symmetric in (a,b,c) and vanishes on the diagonals, so the a<b<c sum is
    # one sixth of the full-cube sum (avoids building index masks)
    return float(np.sum(W * (W + V) / D)) / 6.0


class PackedTriples:
    """CPU baseline for config 5 organised like a real run: the b<c-packed slabs vovvP[i] (v x NP) and
    T2P[i] (o x NP) are built once per occupied index and reused by every triple that touches it (in the
    full job each slab serves ~740 triples), X = sum of six BLAS GEMMs, OpenMP energy sweep."""

    def __init__(self, T1, T2, oovv, vovv, ovoo, eo, ev):
        self.T1, self.T2, self.oovv, self.vovv, self.ovoo = T1, T2, np.ascontiguousarray(oovv), vovv, ovoo
        self.eo, self.ev = np.ascontiguousarray(eo), np.ascontiguousarray(ev)
        self.o, self.v = T1.shape
        self.bu, self.cu = np.triu_indices(self.v, 1)
        self._vovvP, self._T2P = {}, {}

    def slabs(self, i):
        if i not in self._vovvP:
            self._vovvP[i] = np.ascontiguousarray(self.vovv[:, i][:, self.bu, self.cu])
            self._T2P[i] = np.ascontiguousarray(self.T2[i][:, self.bu, self.cu])
        return self._vovvP[i], self._T2P[i]

    def prepare(self, occ):
        for i in occ:
            self.slabs(int(i))

    def triple(self, i, j, k):
        from oracle import oracle as _orc
        T2, ovoo = self.T2, self.ovoo
        (Vi, Ti), (Vj, Tj), (Vk, Tk) = self.slabs(i), self.slabs(j), self.slabs(k)
        Xp = T2[j, k] @ Vi
        Xp -= T2[i, k] @ Vj
        Xp -= T2[j, i] @ Vk
        Xp -= ovoo[:, :, j, k].T @ Ti
        Xp += ovoo[:, :, i, k].T @ Tj
        Xp += ovoo[:, :, j, i].T @ Tk
        return _orc.t_energy_packed(Xp, self.T1, self.oovv, self.eo, self.ev, i, j, k)


def t_flops(o, v, ntriples=None):
    """Algorithmic flop count of SURVEY.md section 8d: one M=v, N=v(v-1)/2, K=3(v+o) GEMM per triple."""
    if ntriples is None:
        ntriples = o * (o - 1) * (o - 2) // 6
    return ntriples * 2.0 * v * (v * (v - 1) // 2) * 3 * (v + o)
