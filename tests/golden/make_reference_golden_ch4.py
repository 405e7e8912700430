"""Config 3 (CH4 / cc-pVTZ, RHF) through the REFERENCE'S OWN sources (oracle/_ref, see make_reference_golden.py):
main.cpp's chain with type = "CCSD" -- RHF, MP2(), CCSD() -- which fits in host memory (so_ints 7 GB, <vv||vv> and W_abef
5.5 GB each); only CCSD_T() cannot run there (4 x o^3 v^3 = 136 GB).  Appends "ch4_ccpvtz_rhf" to
tests/golden/reference_energies.json.  Hours of CPU time on the shim's OpenMP GEMM.

Run here (the reference cannot travel to the GPU box):  python tests/golden/make_reference_golden_ch4.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def main():
    ref.build()
    t0 = time.time()
    r = ref.molecule(os.path.join(ROOT, "geom/methane.xyz"), "cc-pvtz", ref="RHF", type_="CCSD", multiplicity=1, maxiter=400,
                     scf_conv=1e-10, cc_conv=1e-12)  # the reference's plain Roothaan SCF does not reach 1e-12 in 200 iterations here
    r["seconds"] = time.time() - t0
    r["e_t"] = None  # CCSD_T() is not run (see the docstring)
    r["scf_conv"], r["cc_conv"], r["maxiter"] = 1e-10, 1e-12, 400
    r["what"] = "reference main.cpp chain with type CCSD (RHF, MP2(), CCSD()); CCSD_T() not run: 4 x o^3 v^3 = 136 GB"
    print(r, flush=True)
    path = os.path.join(ROOT, "tests", "golden", "reference_energies.json")
    d = json.load(open(path))
    d["ch4_ccpvtz_rhf"] = r
    json.dump(d, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
