"""Writes tests/golden/packing_golden.json: the BASELINE sphere-packing generator (seed 12345,
R 12, solid 0.60) at sizes the CPU oracle still finishes in minutes, tau in Z for the pore phase,
solved by the C restatement (oracle/oi_oracle.c: literal flood mask, stored 7-coefficient matrix,
Jacobi-PCG to 1e-11 on the reference's stopping rule).  The GPU suite compares the CUDA path with
these at full 1e-6 tolerance -- sizes between the seconds-scale oracle cases and the
property-only 512^3 / 1024^3 checks.
Run in the build container:  python tests/golden/make_packing_golden.py   (about 10 minutes)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from openimpala_b200 import synth  # noqa: E402  (the generator is product code; the solve below is the oracle's)
from oracle import oi_c  # noqa: E402

out = {"generator": "openimpala_b200.synth.sphere_packing(n, 12345, 12, 0.60)", "direction": 2, "phase": 1,
       "vlo": -1.0, "vhi": 1.0, "eps": 1e-11, "cases": []}
for n in (96, 128, 192, 256):
    ph = synth.sphere_packing(n, 12345, 12, 0.60)
    r = oi_c.tortuosity(ph.astype(np.int32), 1, 2, -1.0, 1.0, eps=1e-11, maxiter=200000)
    out["cases"].append(dict(n=n, sha256=hashlib.sha256(ph.tobytes()).hexdigest(), phase_count=int((ph == 1).sum()),
                             n_active=r["n_active"], tau=r["tau"], deff=r["deff"], flux_in=r["flux_in"],
                             flux_out=r["flux_out"], oracle_iters=r["iters"], oracle_relres=r["relres"],
                             oracle_solve_s=r["solve_s"]))
    print(out["cases"][-1], flush=True)
    json.dump(out, open(os.path.join(HERE, "packing_golden.json"), "w"), indent=1)
