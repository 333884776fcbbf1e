"""BASELINE.json configs[0] and configs[1]: the reference's sample image
(SampleData_2Phase_stack_3d_1bit.tif, 100^3): volume fraction and tau in X, Y, Z for phase 1
on one B200 through the public class, beside the CPU restatement (C/OpenMP oracle, all host
cores) on the same image, with |d tau| / tau against the committed golden values.

    python tools/report_sample.py > profiles/r1_sample_image_report.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from openimpala_b200.effdiff import calculate_Deff_tensor_homogenization  # noqa: E402
from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre, VolumeFraction  # noqa: E402
from oracle import oi_c, oi_numpy as o  # noqa: E402

gold = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_golden.json")))
ph = o.threshold(o.read_tiff_raw(os.path.join(ROOT, "tests", "golden", "SampleData_2Phase_stack_3d_1bit.tif")), 0.5)
out = {"image": "SampleData_2Phase_stack_3d_1bit.tif", "shape": list(ph.shape), "phase_id": 1, "eps": 1e-9, "rows": []}
pc, tc = VolumeFraction(ph, 1).value()
out["volume_fraction"] = {"phase_count": pc, "total": tc, "vf": pc / tc, "golden_phase_count": 398309}
# warm-up (first construction pays context creation and cudaMalloc)
TortuosityHypre(None, None, None, ph, pc / tc, 1, Direction.X, SolverType.FlexGMRES, "", -1.0, 1.0).value()
for d in (Direction.X, Direction.Y, Direction.Z):
    ref = next(c for c in gold["cases"] if c["phase"] == 1 and c["direction"] == int(d))
    t0 = time.perf_counter()
    t = TortuosityHypre(None, None, None, ph, pc / tc, 1, d, SolverType.FlexGMRES, "", -1.0, 1.0)
    tau = t.value()
    wall = time.perf_counter() - t0
    info = t.last_info
    t0 = time.perf_counter()
    cpu = oi_c.tortuosity(ph, 1, int(d), -1.0, 1.0, eps=1e-9)
    cpu_wall = time.perf_counter() - t0
    out["rows"].append({
        "direction": d.name, "tau_gpu": tau, "tau_golden": ref["tau"], "rel_diff_vs_golden": abs(tau - ref["tau"]) / ref["tau"],
        "active_cells": t._n_active, "active_cells_golden": ref["n_active"], "iterations": info.iterations,
        "rel_residual": info.rel_residual, "gpu_solve_ms": info.solve_ms, "gpu_wall_ms_object_to_tau": 1e3 * wall,
        "dof_iter_per_s": ph.size * info.iterations / (1e-3 * info.solve_ms),
        "cpu_restatement": {"tau": cpu["tau"], "iterations": cpu["iters"], "wall_s": cpu_wall, "cores": oi_c.num_threads(),
                            "kind": "port (C/OpenMP Jacobi-PCG, not HYPRE)"},
        "time_ratio_cpu_over_gpu": cpu_wall / wall})
    t.close()
t0 = time.perf_counter()
D, ok, infos = calculate_Deff_tensor_homogenization(ph, 1)
out["homogenization"] = {"deff": D.tolist(), "converged": ok, "iterations": [i["iterations"] for i in infos],
                         "wall_ms_three_correctors": 1e3 * (time.perf_counter() - t0)}
print(json.dumps(out, indent=1))
