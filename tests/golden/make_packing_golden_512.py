"""Writes tests/golden/packing_golden_512.json: BASELINE configs[2] (the sphere-packing generator at 512^3, seed
12345, R 12, solid 0.60; tau in Z for the pore phase) solved by the C restatement (oracle/oi_oracle.c: literal
flood mask, stored 7-coefficient matrix of 7.5 GB, Jacobi-PCG to 1e-11 on the reference's stopping rule).
One-off: about an hour on this container's cores; the wall time is recorded in the file.
Run in the build container:  OMP_NUM_THREADS=6 python tests/golden/make_packing_golden_512.py"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from openimpala_b200 import synth  # noqa: E402  (the generator is product code; the solve below is the oracle's)
from oracle import oi_c  # noqa: E402

n = 512
t0 = time.time()
ph = synth.sphere_packing(n, 12345, 12, 0.60)
r = oi_c.tortuosity(ph.astype(np.int32), 1, 2, -1.0, 1.0, eps=1e-11, maxiter=400000)
out = {"generator": "openimpala_b200.synth.sphere_packing(512, 12345, 12, 0.60)", "direction": 2, "phase": 1,
       "vlo": -1.0, "vhi": 1.0, "eps": 1e-11, "oracle_threads": oi_c.num_threads(),
       "cases": [dict(n=n, sha256=hashlib.sha256(ph.tobytes()).hexdigest(), phase_count=int((ph == 1).sum()),
                      n_active=r["n_active"], tau=r["tau"], deff=r["deff"], flux_in=r["flux_in"],
                      flux_out=r["flux_out"], oracle_iters=r["iters"], oracle_relres=r["relres"],
                      oracle_solve_s=r["solve_s"], wall_s=time.time() - t0)]}
print(out, flush=True)
json.dump(out, open(os.path.join(HERE, "packing_golden_512.json"), "w"), indent=1)
