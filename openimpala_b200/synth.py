"""Synthetic workloads named in BASELINE.json: random overlapping-sphere packings.

The reference ships no generator for them (data/spheres.tif is a fixed 100^3
example), so the packing is specified here completely by (shape, seed, radius,
solid_target): M = ceil(-ln(1 - solid_target) * N / V_ball) solid spheres of
radius R voxels whose centres are drawn uniformly in the box from PCG64(seed),
x then y then z as three length-M integer draws (Boolean model: expected solid
fraction = solid_target).  Because M is fixed up front, any rank can paint just
its z-slab and all slabs agree.  Voxel value 1 = pore, 0 = solid.
"""
from __future__ import annotations

import hashlib
import math

import numpy as np


def sphere_packing_slab(shape, seed: int = 12345, radius: int = 12, solid_target: float = 0.60,
                        z_begin: int = 0, nz_local: int | None = None) -> np.ndarray:
    """uint8 [nz_local, ny, nx] slab of the global packing of `shape` = (nz, ny, nx)."""
    nz, ny, nx = (int(s) for s in shape)
    if nz_local is None:
        nz_local = nz - z_begin
    r = int(radius)
    g = np.arange(-r, r + 1)
    ball = (g[:, None, None] ** 2 + g[None, :, None] ** 2 + g[None, None, :] ** 2) <= r * r
    vball = int(ball.sum())
    m = int(math.ceil(-math.log(1.0 - solid_target) * (nx * ny * nz) / vball))
    rng = np.random.Generator(np.random.PCG64(seed))
    cx = rng.integers(0, nx, size=m)
    cy = rng.integers(0, ny, size=m)
    cz = rng.integers(0, nz, size=m)
    solid = np.zeros((nz_local, ny, nx), dtype=bool)
    z_end = z_begin + nz_local
    sel = np.nonzero((cz + r >= z_begin) & (cz - r < z_end))[0]
    for s in sel:
        x, y, z = int(cx[s]), int(cy[s]), int(cz[s])
        z0, z1 = max(z - r, z_begin), min(z + r + 1, z_end)
        y0, y1 = max(y - r, 0), min(y + r + 1, ny)
        x0, x1 = max(x - r, 0), min(x + r + 1, nx)
        solid[z0 - z_begin:z1 - z_begin, y0:y1, x0:x1] |= ball[z0 - z + r:z1 - z + r,
                                                               y0 - y + r:y1 - y + r,
                                                               x0 - x + r:x1 - x + r]
    return (~solid).astype(np.uint8)


def sphere_packing(n: int, seed: int = 12345, radius: int = 12, solid_target: float = 0.60) -> np.ndarray:
    return sphere_packing_slab((n, n, n), seed, radius, solid_target)


def describe(a: np.ndarray) -> dict:
    return {"shape_zyx": list(a.shape), "porosity": float(a.mean()),
            "sha256": hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()}
