"""Generates tests/golden/reference_energies.json by RUNNING THE REFERENCE ITSELF: the unmodified
sources of /root/reference compiled against the API shims of oracle/ref_shim (oracle/_ref/
libqcp2_ref.so; see oracle/ref.py).  Each entry is what main.cpp:22-88 computes for one of
BASELINE.json's real-molecule configs -- RHF/UHF, MP2(), CCSD(), CCSD_T() -- at full precision,
plus the 12-digit lines the reference binary (oracle/_ref/P2_ref) prints for the same input.
For the UHF config the reference's converged MO coefficients and orbital energies are saved too
(reference_nh2_boundary.npz): the as-written UHF path (swapped spin labels in
transform_to_so) is NOT invariant to the arbitrary signs of the MO vectors, so its energies are
only reproducible from the very C matrices the reference's eigensolver returned.
Config 3 (CH4/cc-pVTZ) is left out: the reference's CCSD_T needs 4 x o^3 v^3 doubles = 136 GB there.

Run here (the reference cannot travel to the GPU box):  python tests/golden/make_reference_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

CONFIGS = {
    "water_sto3g_rhf": dict(filename="geom/water.xyz", basis="sto-3g", ref="RHF", multiplicity=1),
    "water_ccpvdz_rhf": dict(filename="geom/water.xyz", basis="cc-pvdz", ref="RHF", multiplicity=1),
    "nh2_ccpvdz_uhf": dict(filename="geom/nh2.xyz", basis="cc-pvdz", ref="UHF", multiplicity=2),
}


def main():
    ref.build()
    out = {"_generator": "tests/golden/make_reference_golden.py",
           "_what": "energies computed by the reference's own sources (compiled against oracle/ref_shim) on "
                    "BASELINE.json configs 1, 2 and 4 (maxiter 200, scf_conv = cc_conv = 1e-12, as input.json)"}
    for name, c in CONFIGS.items():
        r = ref.molecule(os.path.join(ROOT, c["filename"]), c["basis"], ref=c["ref"], type_="CCSD(T)",
                         multiplicity=c["multiplicity"], maxiter=200, scf_conv=1e-12, cc_conv=1e-12,
                         want_boundary=(c["ref"] == "UHF"))
        if c["ref"] == "UHF":
            np.savez(os.path.join(ROOT, "tests", "golden", "reference_nh2_boundary.npz"), Ca=r["Ca"], Cb=r["Cb"],
                     eps_a=r["eps_a"], eps_b=r["eps_b"])
            for k in ("ao", "Ca", "Cb", "eps_a", "eps_b"):
                del r[k]
        # and the reference binary on the same input, from the repo root like the reference's own input.json
        cfg = dict(filename=c["filename"], type="CCSD(T)", ref=c["ref"], basis=c["basis"], charge=0,
                   multiplicity=c["multiplicity"], maxiter=200, scf_conv=1e-12, cc_conv=1e-12)
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(cfg, f)
        txt = subprocess.run([ref.BIN, f.name], cwd=ROOT, capture_output=True, text=True, check=True).stdout
        os.unlink(f.name)
        keep = [l for l in txt.splitlines() if any(k in l for k in ("Hartree-Fock Energy", "MP2 Energy", "CCSD energy",
                                                                    "(T) energy", "Total energy", "CC energy converged"))]
        r["p2_ref_stdout"] = keep
        out[name] = r
        print(name, r, flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "reference_energies.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
