"""Writes tests/golden/sample_flexgmres_golden.json: the reference's sample
image solved the way the reference itself solves it -- the assembled
non-symmetric system with identity rows kept, FlexGMRES(20), hypre.eps = 1e-9,
hypre.maxiter = 200 (TortuosityHypre.cpp:142-143, 666-688), vlo/vhi = -1/+1
(Diffusion.cpp defaults) -- by oracle/oi_numpy.py:solve_full_flexgmres.
The preconditioner is an incomplete LU standing in for HYPRE's SMG (not
restatable from the reference tree), so iteration counts are NOT HYPRE's;
tau, D_eff, fluxes and the converged flag are what the reference's rule yields.
Phase 1 only (BASELINE configs[0]/[1]): with this stand-in preconditioner the larger
phase-0 system does not reach 1e-9 inside the reference's 200-iteration cap.
Run in the build container:  python tests/golden/make_flexgmres_golden.py  (about 3 minutes)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oi_numpy as o  # noqa: E402

path = os.path.join(HERE, "SampleData_2Phase_stack_3d_1bit.tif")
ph = o.threshold(o.read_tiff_raw(path), 0.5)
out = {"file": os.path.basename(path), "threshold": 0.5, "vlo": -1.0, "vhi": 1.0, "eps": 1e-9,
       "maxiter": 200, "k_dim": 20, "preconditioner": "scipy spilu(drop_tol=1e-5, fill_factor=20)",
       "cases": []}
for pid in (1,):
    for d in range(3):
        r = o.tortuosity(ph, pid, d, -1.0, 1.0, eps=1e-9, method="flexgmres", maxiter=200)
        out["cases"].append(dict(phase=pid, direction=d, n_active=r.n_active, tau=r.tau, deff=r.deff,
                                 flux_in=r.flux_in, flux_out=r.flux_out, iters=r.iters,
                                 relres=r.relres, converged=r.converged))
        print(out["cases"][-1], flush=True)
        json.dump(out, open(os.path.join(HERE, "sample_flexgmres_golden.json"), "w"), indent=1)
