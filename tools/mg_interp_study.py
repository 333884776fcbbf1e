"""Scratch study (CPU, scipy): does a linear (Kwak-type, mask-aware) prolongation
cut the PCG iteration count of the cell-centred V-cycle compared with the
piecewise-constant aggregation transfer?  Coarse operators stay the 7-point
aggregation-Galerkin sums scaled by 1/2 (what the CUDA path stores).

    python tools/mg_interp_study.py 128 [radius]
    python tools/mg_interp_study.py sample
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


def build(A, coords, shape, scale=0.5, min_cells=64, galerkin_lin=False, wnb=0.25):
    """Levels with both transfers: Pc (piecewise constant) and Pl (linear on the
    simplex parent + nearest face neighbour per axis, only across coarse faces
    that carry a coupling; missing weights fall back to the parent)."""
    levels = [dict(A=A)]
    shp = shape
    while levels[-1]["A"].shape[0] > min_cells and max(shp) > 2:
        Af = levels[-1]["A"]
        k, j, i = coords
        ck, cj, ci = k // 2, j // 2, i // 2
        cshp = tuple((s + 1) // 2 for s in shp)
        lin = (ck * cshp[1] + cj) * cshp[2] + ci
        uniq, inv = np.unique(lin, return_inverse=True)
        n_f, n_c = len(lin), len(uniq)
        Pc = sp.csr_matrix((np.ones(n_f), (np.arange(n_f), inv)), shape=(n_f, n_c))
        Ac = (Pc.T @ Af @ Pc).tocsr() * scale
        # linear prolongation
        lut = -np.ones(int(np.prod(cshp)), dtype=np.int64)
        lut[uniq] = np.arange(n_c)
        rows, cols, vals = [np.arange(n_f)], [inv], [np.ones(n_f)]
        Acoo = Ac.tocsr()
        for axis, (f_idx, c_idx, stride, ext) in enumerate(
                [(k, ck, cshp[1] * cshp[2], cshp[0]), (j, cj, cshp[2], cshp[1]), (i, ci, 1, cshp[2])]):
            side = np.where(f_idx % 2 == 1, 1, -1)
            nb_c = c_idx + side
            ok = (nb_c >= 0) & (nb_c < ext)
            nb_lin = lin + side * stride
            nb = np.where(ok, lut[np.where(ok, nb_lin, 0)], -1)
            ok &= nb >= 0
            # coupling present between parent aggregate and that neighbour?
            coup = np.zeros(n_f, dtype=bool)
            sel = np.nonzero(ok)[0]
            coup[sel] = np.asarray(Acoo[inv[sel], nb[sel]]).ravel() != 0.0
            ok &= coup
            sel = np.nonzero(ok)[0]
            rows += [sel, sel]
            cols += [nb[sel], inv[sel]]
            vals += [np.full(len(sel), wnb), np.full(len(sel), -wnb)]
        Pl = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_f, n_c))
        Pl.sum_duplicates()
        if galerkin_lin:
            Ac = (Pl.T @ Af @ Pl).tocsr()
        levels[-1]["Pc"], levels[-1]["Pl"] = Pc, Pl
        levels.append(dict(A=Ac))
        coords = (uniq // (cshp[1] * cshp[2]), (uniq // cshp[2]) % cshp[1], uniq % cshp[2])
        shp = cshp
    for L in levels:
        L["dinv"] = 1.0 / L["A"].diagonal()
    return levels


def vcycle(levels, l, b, cfg):
    L = levels[l]
    if l == len(levels) - 1:
        return smooth_wjac(L, None, b, cfg["cw"])
    x = smooth_wjac(L, None, b, cfg["w"])
    r = b - L["A"] @ x
    P = L[cfg["P"]]
    R = L[cfg["R"]]
    rc = cfg.get("rscale", 1.0) * (R.T @ r)
    ec = vcycle(levels, l + 1, rc, cfg)
    gamma = cfg.get("gamma", 1) if l < cfg.get("wdepth", 99) else 1
    for _ in range(gamma - 1):      # W-cycle: another visit on the coarse residual
        ec = ec + vcycle(levels, l + 1, rc - levels[l + 1]["A"] @ ec, cfg)
    x = x + P @ ec
    return smooth_wjac(L, x, b, cfg["w"][::-1])


def study(phase, phase_id, direction, cfgs, eps=1e-9, vlo=-1.0, vhi=1.0):
    mask = o.activity_mask(phase, phase_id, direction)
    a, rhs, x0 = o.fill_matrix(phase, mask, phase_id, direction, vlo, vhi)
    A = o.assemble_csr(a, phase.shape)
    Auu, bu, unk, xf = o.eliminate_dirichlet(A, rhs, x0, phase.shape, mask, direction)
    bnorm = o.reference_stop_norm(rhs)
    nz, ny, nx = phase.shape
    lin = np.nonzero(unk)[0]
    coords = (lin // (nx * ny), (lin // nx) % ny, lin % nx)
    print(f"unknowns {Auu.shape[0]}  bnorm {bnorm:.4g}", flush=True)
    cache = {}
    for name, cfg in cfgs.items():
        t = time.time()
        key = (cfg.get("scale", 0.5), cfg.get("galerkin_lin", False), cfg.get("wnb", 0.25))
        if key not in cache:
            cache[key] = build(Auu, coords, phase.shape, scale=key[0], galerkin_lin=key[1], wnb=key[2])
        levels = cache[key]
        x, it, hist = pcg(Auu, bu, x0[unk], lambda r: vcycle(levels, 0, r, cfg), eps * bnorm, maxiter=120)
        sweeps = 2 * len(cfg["w"]) * it
        print(f"  {name:44s} levels {len(levels)} iters {it:4d}  fine sweeps {sweeps:4d}  {time.time()-t:.1f}s",
              flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "96"
    lo_tab = {1: 0.4, 2: 0.25, 3: 0.2, 4: 0.15, 5: 0.12, 6: 0.1}
    cw = cheb_weights(8, 0.05)
    cfgs = {}
    for d in (4, 3, 2):
        w = cheb_weights(d, lo_tab[d])
        cfgs[f"const/const   deg {d} (current)"] = dict(P="Pc", R="Pc", w=w, cw=cw)
        cfgs[f"linear/linearT deg {d}"] = dict(P="Pl", R="Pl", w=w, cw=cw)
        cfgs[f"linear/linearT deg {d} scale 1 (rediscr. x2)"] = dict(P="Pl", R="Pl", w=w, cw=cw, scale=1.0)
        cfgs[f"linear/linearT deg {d} true Galerkin"] = dict(P="Pl", R="Pl", w=w, cw=cw, galerkin_lin=True)
    if len(sys.argv) > 3 and sys.argv[3] == "w":
        cfgs = {}
        for d in (4, 3, 2):
            w = cheb_weights(d, lo_tab[d])
            cfgs[f"const/const deg {d} V"] = dict(P="Pc", R="Pc", w=w, cw=cw)
            cfgs[f"const/const deg {d} W first 3 levels"] = dict(P="Pc", R="Pc", w=w, cw=cw, gamma=2, wdepth=3)
            cfgs[f"const/const deg {d} W first 2 levels"] = dict(P="Pc", R="Pc", w=w, cw=cw, gamma=2, wdepth=2)
            cfgs[f"const/const deg {d} W first level"] = dict(P="Pc", R="Pc", w=w, cw=cw, gamma=2, wdepth=1)
    if which == "sample":
        ph = o.threshold(o.read_tiff_raw("/root/reference/data/SampleData_2Phase_stack_3d_1bit.tif"))
        study(ph, 1, 0, cfgs)
    else:
        n = int(which)
        ph = o.sphere_packing(n, radius=int(sys.argv[2]) if len(sys.argv) > 2 else 12).astype(np.int32)
        print("porosity", ph.mean())
        study(ph, 1, 2, cfgs)
