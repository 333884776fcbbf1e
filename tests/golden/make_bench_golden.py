"""Integer facts of the bench.py workloads (tests/golden/bench_golden.json), computed WITHOUT the CUDA path:
phase-1 voxel count by numpy, percolating (active) count and active inlet / outlet plane cells by
scipy.ndimage.label (6-connectivity) -- the closed form of generateActivityMask.  tau is NOT
produced here: for 512^3 it comes from the C oracle (packing_golden_512.json); for sizes the CPU oracle cannot
solve it is the single-GPU result of the same library (kind "self, N=1"), which makes the N>1 lines of the
scaling run a multi-GPU consistency check, not an independent one -- said so in `tau_kind`.
    python tests/golden/make_bench_golden.py 512 1024 [1280 ...]      (1024^3 needs ~12 GB, a few minutes)"""
import json
import os
import sys
import time

import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from openimpala_b200 import synth  # noqa: E402

PATH = os.path.join(HERE, "bench_golden.json")
gold = json.load(open(PATH)) if os.path.exists(PATH) else {}
for arg in sys.argv[1:]:
    n = int(arg)
    t0 = time.time()
    ph = synth.sphere_packing(n, 12345, 12, 0.60)
    pc = int(np.count_nonzero(ph == 1))
    lab, _ = ndimage.label(ph == 1)                      # default structure: 6-connectivity
    both = np.intersect1d(np.unique(lab[0][lab[0] > 0]), np.unique(lab[-1][lab[-1] > 0]))
    keep = np.zeros(int(lab.max()) + 1, dtype=bool)
    keep[both] = True
    n_active = int(keep[lab].sum()) if n <= 640 else sum(int(keep[lab[k]].sum()) for k in range(n))
    n_in, n_out = int(keep[lab[0]].sum()), int(keep[lab[-1]].sum())
    e = gold.setdefault(f"{n}:2", {})
    e.update(phase_cells=pc, active_cells=n_active, n_in=n_in, n_out=n_out,
             source=e.get("source") or "integers: numpy + scipy.ndimage.label (tests/golden/make_bench_golden.py)",
             integers_wall_s=round(time.time() - t0, 1))
    print(n, e, flush=True)
    json.dump(gold, open(PATH, "w"), indent=1)
