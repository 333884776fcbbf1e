"""Debug helper: the flux-polish case of test_flux_gate_returns_nan_like_the_reference under env combinations."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, math
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import numpy as np
from oracle import oi_numpy as o
from openimpala_b200 import capi
ph = o.threshold(o.read_tiff_raw(%r + "/tests/golden/SampleData_2Phase_stack_3d_1bit.tif"), 0.5)
for polish in (0, 1):
    with capi.Solver(ph.shape, 0, 1, -1.0, 1.0, eps=1e-4, flux_polish=polish) as s:
        s.set_phase(ph); s.build_mask(); info = s.solve(); fin, fout, _, _ = s.fluxes()
        mis = abs(abs(fin) - abs(fout)) / (0.5 * (abs(fin) + abs(fout)))
        print("polish", polish, "iters", info.iterations, "relres %%.3e" %% info.rel_residual, "conv", info.converged, "mismatch %%.3e" %% mis, flush=True)
''' % (ROOT, ROOT, ROOT)
for env in ({}, {"OI_GRAPH": "0"}, {"OI_PAIR": "2"}, {"OI_PAIR": "0"}, {"OI_MG_DEG_COARSE": "4"}, {"OI_COARSE_HALF": "0"}, {"OI_TAIL": "0"}):
    print("== env", env, flush=True)
    r = subprocess.run([sys.executable, "-c", CODE], env={**os.environ, **env}, capture_output=True, text=True)
    print(r.stdout + r.stderr[-800:], flush=True)
