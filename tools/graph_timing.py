"""Stream path against CUDA-graph replay of the PCG iterations (OI_GRAPH=0|1), without and
with the one-CTA coarse tail (OI_TAIL=0|1), on the
reference's sample image and on small sphere packings: solve milliseconds (CUDA events
inside oi_solve) and wall-clock object-to-tau milliseconds, best of a few repetitions.
Run on the GPU box:  python tools/graph_timing.py > gpurun_out/graph_timing.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from openimpala_b200 import capi, synth  # noqa: E402


def read_sample_tiff():
    from PIL import Image
    im = Image.open(os.path.join(ROOT, "tests", "golden", "SampleData_2Phase_stack_3d_1bit.tif"))
    planes = []
    for k in range(im.n_frames):
        im.seek(k)
        planes.append((np.array(im.convert("L")) > 0).astype(np.uint8))
    return np.stack(planes)


def run(ph, direction, reps=5):
    rows = {}
    for mode, tail in (("0", "0"), ("1", "0"), ("1", "1")):
        os.environ["OI_GRAPH"] = mode
        os.environ["OI_TAIL"] = tail
        best_solve, best_wall, its, replays, nodes = 1e30, 1e30, 0, 0, 0
        for _ in range(reps):
            t0 = time.perf_counter()
            with capi.Solver(ph.shape, direction, 1, -1.0, 1.0) as s:
                s.set_phase(ph)
                s.build_mask()
                info = s.solve()
                s.fluxes()
                wall = (time.perf_counter() - t0) * 1e3
                replays, nodes = s.graph_info()
            best_solve = min(best_solve, info.solve_ms)
            best_wall = min(best_wall, wall)
            its = info.iterations
        rows[("graph+tail" if tail == "1" else "graph") if mode == "1" else "stream"] = dict(solve_ms=best_solve, wall_ms=best_wall, iterations=its,
                                                          replays=replays, kernel_nodes=nodes)
    rows["solve_speedup"] = rows["stream"]["solve_ms"] / rows["graph"]["solve_ms"]
    rows["solve_speedup_with_tail"] = rows["stream"]["solve_ms"] / rows["graph+tail"]["solve_ms"]
    return rows


if __name__ == "__main__":
    out = {}
    ph = read_sample_tiff()
    for d, name in enumerate("XYZ"):
        out[f"sample_100^3_{name}"] = run(ph, d)
    sizes = [int(v) for v in os.environ.get("OI_TIMING_SIZES", "64,128,192,256,320").split(",") if v]
    for n in sizes:
        out[f"packing_{n}^3_Z"] = run(synth.sphere_packing(n, 12345, 12, 0.60), 2, reps=3)
    print(json.dumps(out, indent=1))
