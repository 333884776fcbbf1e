"""Scratch study (CPU, scipy): anisotropic cells.  Full 2x2x2 coarsening with one scalar
scale versus strength-based semicoarsening (coarsen axis a only while c_a >= theta * max c)
with per-axis scaling 1/f_a of the aggregated couplings.

    python tools/mg_semicoarsen_study.py 64 1 1 5
"""
import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, _ROOT)
from oracle import oi_numpy as o  # noqa: E402
from tools.mg_prototype import cheb_weights, pcg, smooth_wjac  # noqa: E402


def split_by_axis(a, shape, unk):
    """A_uu = Ax + Ay + Az: couplings along each axis plus the diagonal share they put in
    (couplings to Dirichlet cells stay on the diagonal as sinks)."""
    nz, ny, nx = shape
    N = nx * ny * nz
    strides = (0, -1, 1, -nx, nx, -nx * ny, nx * ny)
    m = np.arange(N)
    parts = []
    for axis, slots in enumerate(((1, 2), (3, 4), (5, 6))):
        rows, cols, vals = [], [], []
        diag = np.zeros(N)
        for s in slots:
            v = a[:, s]
            sel = (v != 0.0) & unk
            diag[sel] += -v[sel]
            nb = m + strides[s]
            off = sel & unk[np.clip(nb, 0, N - 1)]
            rows.append(m[off]); cols.append(nb[off]); vals.append(v[off])
        rows.append(m[unk]); cols.append(m[unk]); vals.append(diag[unk])
        A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
        parts.append(A[unk][:, unk].tocsr())
    return parts


def build(parts, coords, shape, c, mode, theta=0.5, min_cells=64):
    levels = [dict(parts=parts, A=(parts[0] + parts[1] + parts[2]).tocsr())]
    shp = list(shape)            # (nz, ny, nx)
    c = list(c)                  # (cx, cy, cz)
    while levels[-1]["A"].shape[0] > min_cells and max(shp) > 2:
        dims = (shp[2], shp[1], shp[0])                      # nx, ny, nz
        can = [d >= 3 for d in dims]
        if mode == "full":
            f = [2 if ok else 1 for ok in can]
            scale = [0.5, 0.5, 0.5]
        else:
            cmax = max(ca for ca, ok in zip(c, can) if ok) if any(can) else 0.0
            f = [2 if (ok and ca >= theta * cmax) else 1 for ca, ok in zip(c, can)]
            scale = [1.0 / fa for fa in f]
        if f == [1, 1, 1]:
            break
        k, j, i = coords
        ck, cj, ci = k // f[2], j // f[1], i // f[0]
        cshp = [(shp[0] + f[2] - 1) // f[2], (shp[1] + f[1] - 1) // f[1], (shp[2] + f[0] - 1) // f[0]]
        lin = (ck * cshp[1] + cj) * cshp[2] + ci
        uniq, inv = np.unique(lin, return_inverse=True)
        P = sp.csr_matrix((np.ones(len(lin)), (np.arange(len(lin)), inv)), shape=(len(lin), len(uniq)))
        cparts = [(P.T @ levels[-1]["parts"][ax] @ P).tocsr() * scale[ax] for ax in range(3)]
        levels[-1]["P"] = P
        levels.append(dict(parts=cparts, A=(cparts[0] + cparts[1] + cparts[2]).tocsr()))
        others = [f[1] * f[2], f[0] * f[2], f[0] * f[1]]
        c = [c[ax] * others[ax] * scale[ax] for ax in range(3)]
        coords = (uniq // (cshp[1] * cshp[2]), (uniq // cshp[2]) % cshp[1], uniq % cshp[2])
        shp = cshp
        levels[-1]["f"] = f
    for L in levels:
        L["dinv"] = 1.0 / L["A"].diagonal()
    return levels


def vcycle(levels, l, b, w, cw):
    L = levels[l]
    if l == len(levels) - 1:
        return smooth_wjac(L, None, b, cw)
    x = smooth_wjac(L, None, b, w)
    ec = vcycle(levels, l + 1, L["P"].T @ (b - L["A"] @ x), w, cw)
    x = x + L["P"] @ ec
    return smooth_wjac(L, x, b, w[::-1])


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dx = tuple(float(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (1.0, 1.0, 5.0)
    from scipy import ndimage
    rng = np.random.default_rng(3)
    fld = ndimage.gaussian_filter(rng.standard_normal((n, n, n)), 2.0)
    ph = (fld > np.quantile(fld, 0.5)).astype(np.int32)
    for direction in (2, 0):
        mask = o.activity_mask(ph, 1, direction)
        a, rhs, x0 = o.fill_matrix(ph, mask, 1, direction, -1.0, 1.0, dx=dx)
        A = o.assemble_csr(a, ph.shape)
        Auu, bu, unk, xf = o.eliminate_dirichlet(A, rhs, x0, ph.shape, mask, direction)
        parts = split_by_axis(a, ph.shape, unk)
        assert abs((parts[0] + parts[1] + parts[2]) - Auu).max() < 1e-12
        bnorm = o.reference_stop_norm(rhs)
        lin = np.nonzero(unk)[0]
        coords = (lin // (n * n), (lin // n) % n, lin % n)
        c = (1 / dx[0] ** 2, 1 / dx[1] ** 2, 1 / dx[2] ** 2)
        w, cw = cheb_weights(4, 0.15), cheb_weights(8, 0.05)
        for mode, theta in (("full", 0), ("semi", 0.5), ("semi", 0.3), ("semi", 0.7)):
            t = time.time()
            lv = build(parts, coords, ph.shape, c, mode, theta)
            x, it, hist = pcg(Auu, bu, x0[unk], lambda r: vcycle(lv, 0, r, w, cw), 1e-9 * bnorm, maxiter=300)
            print(f"n={n} dx={dx} dir={direction} {mode} theta={theta}: levels {len(lv)} factors "
                  f"{[L.get('f') for L in lv[1:4]]} iters {it}  {time.time() - t:.1f}s", flush=True)
