"""GPU sweep over geometries: iteration counts / convergence of the MG-PCG solve away from the
benchmark packing (low porosity near the percolation threshold, blobby random fields, thin
channels, anisotropic cells, non-cubic boxes)."""
import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys
import time

import numpy as np

sys.path.insert(0, _ROOT)
from openimpala_b200 import capi, synth  # noqa: E402
from openimpala_b200.tortuosity import tau_from_fluxes  # noqa: E402


def blobs(shape, seed, porosity, sigma):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    f = ndimage.gaussian_filter(rng.standard_normal(shape).astype(np.float32), sigma)
    return (f > np.quantile(f, 1.0 - porosity)).astype(np.uint8)


def run(name, ph, direction=2, dx=(1.0, 1.0, 1.0), problem=capi.OI_PROBLEM_TORTUOSITY, maxiter=200):
    t0 = time.time()
    with capi.Solver(ph.shape, direction, 1, -1.0, 1.0, dx=dx, problem=problem, maxiter=maxiter) as s:
        s.set_phase(ph)
        n_active = s.build_mask()
        if n_active == 0:
            print(f"{name:46s} no percolating cells", flush=True)
            return
        info = s.solve()
        extra = ""
        if problem == capi.OI_PROBLEM_TORTUOSITY:
            fin, fout, _, _ = s.fluxes()
            n = ph.shape[2 - direction]
            ext = [ph.shape[2] * dx[0], ph.shape[1] * dx[1], ph.shape[0] * dx[2]]
            area = ext[0] * ext[1] * ext[2] / ext[direction]
            tau, _, cons = tau_from_fluxes(fin, fout, n_active / ph.size, ext[direction], area, -1.0, 1.0)
            extra = f"tau {tau:.6f} flux mismatch {abs(abs(fin) - abs(fout)) / max(abs(fin), 1e-300):.1e}"
        print(f"{name:46s} active {n_active / ph.size:.3f} iters {info.iterations:4d} relres {info.rel_residual:.2e} "
              f"conv {info.converged} solve {info.solve_ms:8.1f} ms {extra}", flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for solid in (0.4, 0.6, 0.7, 0.75):
        run(f"packing {n}^3 solid {solid}", synth.sphere_packing(n, 12345, 12, solid))
    run(f"packing {n}^3 R=4 solid 0.6", synth.sphere_packing(n, 7, 4, 0.6))
    for por, sig in ((0.5, 2.0), (0.35, 2.0), (0.3, 4.0), (0.25, 1.0), (0.6, 6.0)):
        run(f"blobs {n}^3 porosity {por} sigma {sig}", blobs((n, n, n), 3, por, sig))
    run(f"blobs {n}x{n // 2}x{n // 4} X", blobs((n // 4, n // 2, n), 5, 0.5, 2.0), direction=0)
    run(f"blobs {n}^3 dx=(1,1,2)", blobs((n, n, n), 3, 0.5, 2.0), dx=(1.0, 1.0, 2.0))
    run(f"blobs {n}^3 dx=(1,1,5)", blobs((n, n, n), 3, 0.5, 2.0), dx=(1.0, 1.0, 5.0))
    run(f"blobs {n}^3 dx=(3,1,1) X", blobs((n, n, n), 3, 0.5, 2.0), direction=0, dx=(3.0, 1.0, 1.0))
    run(f"open box {n}^3", np.ones((n, n, n), np.uint8))
    ph = np.zeros((n, n, n), np.uint8)
    ph[:, n // 2, n // 2] = 1
    ph[:, n // 2 - 1:n // 2 + 2, 3] = 1
    run(f"single-voxel channels {n}^3", ph)
    run(f"cell problem blobs {n}^3 por 0.5", blobs((n, n, n), 3, 0.5, 2.0), direction=0, problem=capi.OI_PROBLEM_CELL, maxiter=1000)
    run(f"cell problem packing {n}^3", synth.sphere_packing(n, 12345, 12, 0.6), direction=2, problem=capi.OI_PROBLEM_CELL, maxiter=1000)
