"""Regenerates tests/golden/ch4_ccpvtz_oracle.json: BASELINE.json config 3 (CH4 / cc-pVTZ, RHF) on
the CPU only -- host SCF substrate (scf_conv 1e-10, see tools/run_configs.py) -> numpy/einsum
oracle MP2, CCSD (cc_conv 1e-12, reference update order) and per-triple (T).  Takes ~15 min on 8
cores, which is why the numbers are committed instead of being recomputed in the GPU test."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import fock_blocks, molecule_case  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402

t0 = time.time()
case = molecule_case("methane.xyz", "cc-pvtz", conv=1e-10, maxiter=400)
no, nv = case["no"], case["nv"]
mo = npo.transform_ao_mo(case["ao"], case["C"], case["C"])
so = npo.transform_to_so(mo, mo, mo)
In = npo.get_integrals(so, no)
del mo, so
foo, fov, fvv = fock_blocks(case["eps"], case["eps"], no)
eo, ev = np.diag(foo).copy(), np.diag(fvv).copy()
T2, e2 = npo.mp2(In["oovv"], eo, ev)
t1, t2, hist = npo.ccsd(In, foo, fov, fvv, T2, 200, 1e-12)
et = npo.ccsd_t_blocked(t1, t2, In["oovv"], In["vovv"], In["ovoo"], eo, ev, True)
json.dump(dict(e_scf=case["e_scf"], e_mp2=float(e2), e_ccsd=float(hist[-1]), iters=len(hist) - 1,
               e_hist=list(map(float, hist)), e_t=float(et), seconds=time.time() - t0),
          open(os.path.join(HERE, "ch4_ccpvtz_oracle.json"), "w"), indent=1)
