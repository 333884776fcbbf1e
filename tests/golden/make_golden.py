"""Regenerates tests/golden/sample_golden.json from the numpy/scipy oracle
(oracle/oi_numpy.py) on the reference's sample image.  Run in the build
container:  python tests/golden/make_golden.py   (about 3 minutes)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oi_numpy as o  # noqa: E402

path = os.path.join(HERE, "SampleData_2Phase_stack_3d_1bit.tif")
ph = o.threshold(o.read_tiff_raw(path), 0.5)
out = {"file": os.path.basename(path), "md5": hashlib.md5(open(path, "rb").read()).hexdigest(),
       "shape_zyx": list(ph.shape), "threshold": 0.5, "vlo": -1.0, "vhi": 1.0, "eps": 1e-12,
       "phase_count": {}, "cases": []}
for pid in (0, 1):
    out["phase_count"][str(pid)] = o.volume_fraction_counts(ph, pid)[0]
    for d in range(3):
        r = o.tortuosity(ph, pid, d, -1.0, 1.0, eps=1e-12)
        m = o.activity_mask(ph, pid, d)
        n = ph.shape[2 - d]
        out["cases"].append(dict(phase=pid, direction=d, n_active=r.n_active, active_vf=r.active_vf,
                                 n_in=int(o._plane(m, d, 0).sum()), n_out=int(o._plane(m, d, n - 1).sum()),
                                 tau=r.tau, deff=r.deff, flux_in=r.flux_in, flux_out=r.flux_out,
                                 oracle_iters=r.iters, oracle_relres=r.relres,
                                 mask_sha256=hashlib.sha256(m.astype("u1").tobytes()).hexdigest()))
        print(out["cases"][-1], flush=True)
json.dump(out, open(os.path.join(HERE, "sample_golden.json"), "w"), indent=1)
