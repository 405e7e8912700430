"""GPU parity against THE REFERENCE ITSELF (run with -m gpu): every C-ABI entry point vs the function of the
same name in oracle/_ref/libqcp2_ref.so -- the reference's own, unmodified src/{integrals,mp2,cc}.cpp compiled
against the API shims of oracle/ref_shim (built where /root/reference exists; the prebuilt .so files travel to
the GPU box with the snapshot).  Same seeded inputs on both sides, no oracle restatement in between.
Integer/gather work bit-exact, tensors <= 1e-12 relative, energies <= 1e-10 Eh, iteration counts equal."""
import numpy as np
import pytest
from helpers import fock_blocks, rhf_case, uhf_case

import qcp2_b200
from oracle import ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref was not built (no /root/reference at build time)")]
E_TOL = 1e-10

CASES = {
    "rhf_n6": lambda: rhf_case(6, 2, 61),
    "rhf_n7_odd_pairs": lambda: rhf_case(7, 3, 62),
    "uhf_n5_3a2b": lambda: uhf_case(5, 3, 2, 63),
}


def close(a, b, rel=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    return (float(np.max(np.abs(a - b))) <= rel * scale) if b.size else True


@pytest.fixture(scope="module", params=sorted(CASES))
def staged(request):
    """Boundary inputs and every intermediate stage, computed by the REFERENCE's functions."""
    case = CASES[request.param]()
    uhf = "Ca" in case
    no, nv = case["no"], case["nv"]
    if uhf:
        aa = ref.transform_ao_mo(case["ao"], case["Ca"], case["Ca"])
        bb = ref.transform_ao_mo(case["ao"], case["Cb"], case["Cb"])
        ab = ref.transform_ao_mo(case["ao"], case["Ca"], case["Cb"])
        ea, eb = case["eps_a"], case["eps_b"]
    else:
        aa = bb = ab = ref.transform_ao_mo(case["ao"], case["C"], case["C"])
        ea = eb = case["eps"]
    so = ref.transform_to_so(aa, bb, ab)
    F = ref.make_so_moes(ea, eb)
    I = ref.get_integrals(so, no, nv)
    den = ref.make_denom(F, no, nv)
    T2g, e_mp2 = ref.run_mp2(I["oovv"], den)
    foo, fov, fvv = fock_blocks(ea, eb, no)
    return dict(case=case, uhf=uhf, aa=aa, bb=bb, ab=ab, so=so, F=F, I=I, den=den, T2g=T2g, e_mp2=e_mp2, foo=foo, fov=fov, fvv=fvv,
                ea=ea, eb=eb, whole=ref.post_hf(case, maxiter=80, conv=1e-12))


def amplitudes(o, v, seed):
    rng = np.random.default_rng(seed)
    return 0.1 * rng.standard_normal((o, v)), 0.05 * rng.standard_normal((o, o, v, v))


def test_integral_functions(ctx, staged):
    s, case = staged, staged["case"]
    no, nv = case["no"], case["nv"]
    C1, C2 = (case["Ca"], case["Cb"]) if s["uhf"] else (case["C"], case["C"][:, ::-1].copy())
    assert close(ctx.transform_ao_mo(case["ao"], C1, C2), ref.transform_ao_mo(case["ao"], C1, C2))     # integrals.cpp:95-134
    assert np.array_equal(ctx.transform_to_so(s["aa"], s["bb"], s["ab"]), s["so"])                     # :136-167, bit-exact
    for spec in qcp2_b200.INT_NAMES:
        assert np.array_equal(ctx.slice_ints(s["so"], no, nv, spec), s["I"][spec]), spec                # :184-207
    assert np.array_equal(ctx.make_so_moes(s["ea"], s["eb"]), s["F"])                                  # mp2.cpp:11-21
    assert np.array_equal(ctx.make_denom(s["F"], no, nv), s["den"])                                    # integrals.cpp:169-181
    T, e = ctx.run_mp2(s["I"]["oovv"], s["den"])                                                       # mp2.cpp:23-42
    assert close(T, s["T2g"], 1e-15) and abs(e - s["e_mp2"]) < 1e-14
    assert np.array_equal(ctx.make_fock(s["ea"], s["eb"], 0, no), ref.make_fock(s["ea"], s["eb"], 0, no))   # cc.cpp:25-34
    assert np.array_equal(ctx.make_D_ia(s["foo"], s["fvv"]), ref.make_D(s["foo"], s["fvv"], 0))        # cc.cpp:73-83
    assert np.array_equal(ctx.make_D_ijab(s["foo"], s["fvv"]), ref.make_D(s["foo"], s["fvv"], 1))      # :86-100
    assert close(ctx.make_D_triples(s["foo"], s["fvv"]), ref.make_D(s["foo"], s["fvv"], 2), 1e-15)     # :102-123


def test_cc_functions(ctx, staged):
    s, case = staged, staged["case"]
    no, nv = case["no"], case["nv"]
    Ts, Td = amplitudes(no, nv, 71)        # unsymmetric: nothing may lean on permutational symmetry
    fov = 0.01 * np.random.default_rng(9).standard_normal((no, nv))
    I, foo, fvv = s["I"], s["foo"], s["fvv"]
    assert close(ctx.make_tau(Ts, Td), ref.make_tau(Ts, Td), 1e-15)                                    # cc.cpp:125-139
    assert close(ctx.make_tau_bar(Ts, Td), ref.make_tau(Ts, Td, True), 1e-15)                          # :141-156
    assert close(ctx.multiply_Ts(Ts), ref.multiply_Ts(Ts), 1e-15)                                      # :158-172
    Xr = ref.update_intermediates(Ts, Td, I, foo, fov, fvv)                                            # :175-255
    X = ctx.update_intermediates(Ts, Td, I, foo, fov, fvv)
    for k in qcp2_b200.INTER_NAMES:
        assert close(X[k], Xr[k]), k
    assert close(ctx.make_T1(Ts, Td, I, Xr, foo, fov, fvv), ref.make_T1(Ts, Td, I, Xr, foo, fov, fvv))   # :257-292
    assert close(ctx.make_T2(Ts, Td, I, Xr, foo, fvv), ref.make_T2(Ts, Td, I, Xr, foo, fvv))             # :294-387
    assert abs(ctx.ccsd_energy(Ts, Td, I["oovv"], fov) - ref.ccsd_energy(Ts, Td, I["oovv"], fov)) < 1e-13   # :390-405
    eo, ev = np.diag(foo).copy(), np.diag(fvv).copy()
    want = ref.ccsd_t(Ts, Td, I["oovv"], I["vovv"], I["ovoo"], foo, fvv)                               # :454-560, literal
    got = ctx.CCSD_T(Ts, Td, I["oovv"], I["vovv"], I["ovoo"], eo, ev)                                  # AUTO -> LITERAL here
    assert abs(got - want) <= 1e-12 * max(1.0, abs(want))


def test_mp2_ccsd_t_chain(ctx, staged):
    """MP2() body -> CCSD() -> make_moe_tensors -> CCSD_T() of the reference vs the stepwise and the fused
    entry points: energies, Jacobi iteration count (update order Q1, fabs test Q3), amplitudes."""
    s, case, w = staged, staged["case"], staged["whole"]
    uhf, no, nv = s["uhf"], case["no"], case["nv"]
    Ca = case["Ca"] if uhf else case["C"]
    e_mp2, so, T2g = ctx.MP2(case["ao"], Ca, case.get("Cb"), s["ea"], s["eb"], no, nv, uhf=uhf)
    assert abs(e_mp2 - w["e_mp2"]) < E_TOL and close(T2g, w["T2_mp2"])
    res = ctx.CCSD(s["foo"], None if uhf else s["fov"], s["fvv"], T2g, 80, 1e-12, so_ints=so, want_sliced=True)
    assert res["iters"] == w["iters"]
    assert abs(res["ccsd_energy"] - w["e_ccsd"]) < E_TOL
    assert close(res["T1"], w["T1"], 1e-11) and close(res["T2"], w["T2"], 1e-11)
    S = res["sliced_ints"]
    e_t = ctx.CCSD_T(res["T1"], res["T2"], S["oovv"], S["vovv"], S["ovoo"], np.diag(s["foo"]).copy(), np.diag(s["fvv"]).copy())
    assert abs(e_t - w["e_t"]) < E_TOL
    r = ctx.post_hf(case["ao"], Ca, case.get("Cb"), s["ea"], s["eb"] if uhf else None, no, nv, uhf=uhf, stages=3, maxiter=80,
                    cc_conv=1e-12)
    assert abs(r["e_mp2"] - w["e_mp2"]) < E_TOL and abs(r["e_ccsd"] - w["e_ccsd"]) < E_TOL and abs(r["e_t"] - w["e_t"]) < E_TOL
    assert r["iters"] == w["iters"]
