"""Regenerates tests/golden/energies.json.

The reference cannot be built here and ships no golden vectors, so these
fixtures pin the two in-repo oracles against each other and against drift: every number is
produced by the literal C++ oracle (oracle/oracle.cpp) on seeded pseudo-molecules
(tests/helpers.py) and cross-checked with the independent numpy/einsum oracle before it is
written.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import oracle_pipeline, rhf_case, uhf_case  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402

CASES = {
    "rhf_n5_docc2_seed11": dict(kind="rhf", n=5, ndocc=2, seed=11),
    "rhf_n7_docc3_seed5": dict(kind="rhf", n=7, ndocc=3, seed=5),
    "uhf_n5_a3_b2_seed3": dict(kind="uhf", n=5, na=3, nb=2, seed=3),
}


def build(spec):
    if spec["kind"] == "rhf":
        return rhf_case(spec["n"], spec["ndocc"], spec["seed"])
    return uhf_case(spec["n"], spec["na"], spec["nb"], spec["seed"])


def main():
    out = {}
    for name, spec in CASES.items():
        case = build(spec)
        r = oracle_pipeline(case)
        eo, ev = np.diag(r["foo"]).copy(), np.diag(r["fvv"]).copy()
        # cross-oracle: (T) by the per-triple restatement, no symmetry assumed
        et_np = npo.ccsd_t_blocked(r["T1"], r["T2"], r["ints"]["oovv"], r["ints"]["vovv"], r["ints"]["ovoo"], eo, ev,
                                   False)
        assert abs(et_np - r["e_t"]) < 1e-13, (name, et_np, r["e_t"])
        if spec["kind"] == "rhf":  # einsum CCSD is only valid for ket-antisymmetric integrals
            In = npo.get_integrals(r["so"], case["no"])
            _, _, hist = npo.ccsd(In, r["foo"], r["fov"], r["fvv"], r["T2_mp2"], 100, 1e-12)
            assert abs(hist[-1] - r["e_ccsd"]) < 1e-13, (name, hist[-1], r["e_ccsd"])
        out[name] = dict(spec=spec, no=case["no"], nv=case["nv"], e_mp2=r["e_mp2"], e_ccsd=r["e_ccsd"],
                         iters=r["iters"], e_hist=list(map(float, r["e_hist"])), e_t=r["e_t"],
                         so_checksum=float(np.sum(r["so"] * np.arange(1, r["so"].size + 1).reshape(r["so"].shape))),
                         t2_checksum=float(np.sum(np.abs(r["T2"]))))
        print(name, out[name]["e_mp2"], out[name]["e_ccsd"], out[name]["iters"], out[name]["e_t"])
    with open(os.path.join(HERE, "energies.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
