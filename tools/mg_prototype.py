"""Scratch study (CPU, scipy): how many PCG iterations does a cell-centred
aggregation-Galerkin V-cycle need on binary porous geometry?  Drives the
design choices in DESIGN.md (smoother, sweeps, coarse scaling)."""
import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, _ROOT)
from oracle import oi_numpy as o  # noqa: E402


def build_hierarchy(Auu_shape_info, A, idx3, shape, min_cells=64, scale=0.5):
    """A: SPD fine operator on unknowns; idx3: (k,j,i) coords of each unknown.
    Aggregates = 2x2x2 geometric blocks containing >=1 unknown."""
    levels = [dict(A=A)]
    coords = idx3
    shp = shape
    while levels[-1]["A"].shape[0] > min_cells and max(shp) > 2:
        ck, cj, ci = coords[0] // 2, coords[1] // 2, coords[2] // 2
        cshp = tuple((s + 1) // 2 for s in shp)
        lin = (ck * cshp[1] + cj) * cshp[2] + ci
        uniq, inv = np.unique(lin, return_inverse=True)
        n_f, n_c = len(lin), len(uniq)
        P = sp.csr_matrix((np.ones(n_f), (np.arange(n_f), inv)), shape=(n_f, n_c))
        Ac = (P.T @ levels[-1]["A"] @ P).tocsr() * scale
        levels[-1]["P"] = P
        levels.append(dict(A=Ac))
        ck2 = uniq // (cshp[1] * cshp[2])
        cj2 = (uniq // cshp[2]) % cshp[1]
        ci2 = uniq % cshp[2]
        coords = (ck2, cj2, ci2)
        shp = cshp
    for L in levels:
        L["dinv"] = 1.0 / L["A"].diagonal()
    return levels


def smooth_jacobi(L, x, b, n, omega):
    A, dinv = L["A"], L["dinv"]
    for _ in range(n):
        if x is None:
            x = omega * dinv * b
        else:
            x = x + omega * dinv * (b - A @ x)
    return x


def smooth_cheby(L, x, b, degree, lo_frac=0.25, lmax=2.0):
    """Chebyshev on D^-1 A over [lmax*lo_frac, lmax]; symmetric polynomial."""
    A, dinv = L["A"], L["dinv"]
    lmin = lmax * lo_frac
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    r = b if x is None else b - A @ x
    d = dinv * r / theta
    x = d if x is None else x + d
    for _ in range(degree - 1):
        rho_new = 1.0 / (2.0 * sigma - rho)
        r = b - A @ x
        d = rho_new * rho * d + 2.0 * rho_new / delta * (dinv * r)
        x = x + d
        rho = rho_new
    return x


def vcycle(levels, l, b, cfg):
    L = levels[l]
    if l == len(levels) - 1:
        return smooth_jacobi(L, None, b, cfg["coarse_sweeps"], cfg["omega"])
    if cfg["smoother"] == "jacobi":
        x = smooth_jacobi(L, None, b, cfg["nu"], cfg["omega"])
    else:
        x = smooth_cheby(L, None, b, cfg["nu"], cfg["lo_frac"])
    r = b - L["A"] @ x
    rc = L["P"].T @ r
    ec = vcycle(levels, l + 1, rc, cfg)
    x = x + L["P"] @ ec
    if cfg["smoother"] == "jacobi":
        x = smooth_jacobi(L, x, b, cfg["nu"], cfg["omega"])
    else:
        x = smooth_cheby(L, x, b, cfg["nu"], cfg["lo_frac"])
    return x


def pcg(A, b, x0, prec, tol_abs, maxiter=500):
    x = x0.copy()
    r = b - A @ x
    z = prec(r)
    p = z.copy()
    rz = r @ z
    hist = [np.linalg.norm(r)]
    for it in range(1, maxiter + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        rn = np.linalg.norm(r)
        hist.append(rn)
        if rn <= tol_abs:
            return x, it, hist
        z = prec(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxiter, hist


def study(phase, phase_id, direction, cfgs, eps=1e-9, vlo=-1.0, vhi=1.0):
    mask = o.activity_mask(phase, phase_id, direction)
    a, rhs, x0 = o.fill_matrix(phase, mask, phase_id, direction, vlo, vhi)
    A = o.assemble_csr(a, phase.shape)
    Auu, bu, unk, xf = o.eliminate_dirichlet(A, rhs, x0, phase.shape, mask, direction)
    bnorm = o.reference_stop_norm(rhs)
    nz, ny, nx = phase.shape
    lin = np.nonzero(unk)[0]
    idx3 = (lin // (nx * ny), (lin // nx) % ny, lin % nx)
    print(f"unknowns {Auu.shape[0]}  bnorm {bnorm:.4g}")
    out = {}
    for name, cfg in cfgs.items():
        t = time.time()
        levels = build_hierarchy(None, Auu, idx3, phase.shape, scale=cfg.get("scale", 0.5))
        prec = (lambda r: vcycle(levels, 0, r, cfg)) if cfg["smoother"] != "none" else (lambda r: r / Auu.diagonal())
        x, it, hist = pcg(Auu, bu, x0[unk], prec, eps * bnorm, maxiter=cfg.get("maxiter", 400))
        xfull = xf.copy()
        xfull[unk] = x
        fin, fout, _, _ = o.global_fluxes(xfull.reshape(phase.shape), mask, direction)
        tau, deff = o.tau_from_fluxes(fin, fout, mask.sum() / phase.size, phase.shape, direction, vlo, vhi)
        print(f"  {name:28s} levels {len(levels)} iters {it:4d}  tau {tau:.10f}  "
              f"fluxmis {abs(abs(fin)-abs(fout))/abs(fin):.2e}  {time.time()-t:.1f}s")
        out[name] = it
    return out


if __name__ == "__main__" and not (len(sys.argv) > 3):
    which = sys.argv[1] if len(sys.argv) > 1 else "sample"
    cfgs = {
        "V(1,1) jac w=0.8 s=.5": dict(smoother="jacobi", nu=1, omega=0.8, coarse_sweeps=20, scale=0.5),
        "V(2,2) jac w=0.8 s=.5": dict(smoother="jacobi", nu=2, omega=0.8, coarse_sweeps=20, scale=0.5),
        "V(2,2) jac w=0.8 s=1": dict(smoother="jacobi", nu=2, omega=0.8, coarse_sweeps=20, scale=1.0),
        "V cheb2 lo=.25 s=.5": dict(smoother="cheby", nu=2, omega=0.8, lo_frac=0.25, coarse_sweeps=20, scale=0.5),
        "V cheb3 lo=.2 s=.5": dict(smoother="cheby", nu=3, omega=0.8, lo_frac=0.2, coarse_sweeps=20, scale=0.5),
        "V cheb4 lo=.15 s=.5": dict(smoother="cheby", nu=4, omega=0.8, lo_frac=0.15, coarse_sweeps=20, scale=0.5),
    }
    if which == "sample":
        ph = o.threshold(o.read_tiff_raw("/root/reference/data/SampleData_2Phase_stack_3d_1bit.tif"))
        study(ph, 1, 0, cfgs)
        study(ph, 0, 2, cfgs)
    else:
        n = int(which)
        ph = o.sphere_packing(n, radius=int(sys.argv[2]) if len(sys.argv) > 2 else 12).astype(np.int32)
        print("porosity", ph.mean())
        study(ph, 1, 2, cfgs)


def cheb_weights(degree, lo_frac, lmax=2.0):
    """Jacobi weights w_k = 1/root_k of the degree-n Chebyshev polynomial on
    [lo_frac*lmax, lmax]: same polynomial as smooth_cheby, no momentum vector."""
    a, b = lo_frac * lmax, lmax
    k = np.arange(1, degree + 1)
    roots = 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * (2 * k - 1) / (2 * degree))
    return 1.0 / roots


def smooth_wjac(L, x, b, weights):
    A, dinv = L["A"], L["dinv"]
    for w in weights:
        if x is None:
            x = w * dinv * b
        else:
            x = x + w * dinv * (b - A @ x)
    return x


def vcycle2(levels, l, b, cfg):
    L = levels[l]
    if l == len(levels) - 1:
        return smooth_wjac(L, None, b, cfg["cw"])
    x = smooth_wjac(L, None, b, cfg["w"])
    r = b - L["A"] @ x
    rc = L["P"].T @ r
    ec = vcycle2(levels, l + 1, rc, cfg)
    for _ in range(cfg.get("gamma", 1) - 1):   # W-cycle: re-solve coarse residual
        ec = ec + vcycle2(levels, l + 1, rc - levels[l + 1]["A"] @ ec, cfg)
    x = x + L["P"] @ ec
    return smooth_wjac(L, x, b, cfg["w"][::-1])


def study2(phase, phase_id, direction, cfgs, eps=1e-9, vlo=-1.0, vhi=1.0):
    mask = o.activity_mask(phase, phase_id, direction)
    a, rhs, x0 = o.fill_matrix(phase, mask, phase_id, direction, vlo, vhi)
    A = o.assemble_csr(a, phase.shape)
    Auu, bu, unk, xf = o.eliminate_dirichlet(A, rhs, x0, phase.shape, mask, direction)
    bnorm = o.reference_stop_norm(rhs)
    nz, ny, nx = phase.shape
    lin = np.nonzero(unk)[0]
    idx3 = (lin // (nx * ny), (lin // nx) % ny, lin % nx)
    print(f"unknowns {Auu.shape[0]}  bnorm {bnorm:.4g}")
    cache = {}
    for name, cfg in cfgs.items():
        t = time.time()
        key = (cfg.get("scale", 0.5), cfg.get("min_cells", 64))
        if key not in cache:
            cache[key] = build_hierarchy(None, Auu, idx3, phase.shape, scale=key[0], min_cells=key[1])
        levels = cache[key]
        x, it, hist = pcg(Auu, bu, x0[unk], lambda r: vcycle2(levels, 0, r, cfg), eps * bnorm, maxiter=400)
        print(f"  {name:34s} levels {len(levels)} iters {it:4d} {time.time()-t:.1f}s")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "w":
    cf = {}
    for deg, lo in ((2, .25), (3, .2), (3, .3), (4, .15), (4, .25), (5, .12)):
        for sc in (0.5, 0.6):
            cf[f"wjac d{deg} lo{lo} s{sc}"] = dict(w=cheb_weights(deg, lo), cw=cheb_weights(8, 0.05), scale=sc)
    cf["wjac d3 lo.2 s.5 W"] = dict(w=cheb_weights(3, .2), cw=cheb_weights(8, 0.05), scale=.5, gamma=2)
    cf["wjac d4 lo.15 s.5 W"] = dict(w=cheb_weights(4, .15), cw=cheb_weights(8, 0.05), scale=.5, gamma=2)
    cf["wjac d2 lo.25 s.5 W"] = dict(w=cheb_weights(2, .25), cw=cheb_weights(8, 0.05), scale=.5, gamma=2)
    if sys.argv[1] == "sample":
        ph = o.threshold(o.read_tiff_raw("/root/reference/data/SampleData_2Phase_stack_3d_1bit.tif"))
        study2(ph, 1, 0, cf)
    else:
        ph = o.sphere_packing(int(sys.argv[1]), radius=int(sys.argv[2])).astype(np.int32)
        study2(ph, 1, 2, cf)


def vcycle3(levels, l, b, cfg):
    """gamma=2 only while l < cfg['wdepth'] (W on the big levels, V below)."""
    L = levels[l]
    if l == len(levels) - 1:
        return smooth_wjac(L, None, b, cfg["cw"])
    x = smooth_wjac(L, None, b, cfg["w"])
    r = b - L["A"] @ x
    rc = L["P"].T @ r
    ec = vcycle3(levels, l + 1, rc, cfg)
    if l < cfg.get("wdepth", 0):
        ec = ec + vcycle3(levels, l + 1, rc - levels[l + 1]["A"] @ ec, cfg)
    x = x + L["P"] @ ec
    return smooth_wjac(L, x, b, cfg["w"][::-1])


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "w3":
    vcycle2 = vcycle3
    cf = {}
    for deg, lo in ((1, .4), (2, .25), (3, .2)):
        for sc in (0.5, 0.7, 1.0):
            for wd in (0, 2, 3, 99):
                cf[f"d{deg} lo{lo} s{sc} wdepth{wd}"] = dict(w=cheb_weights(deg, lo), cw=cheb_weights(8, 0.05), scale=sc, wdepth=wd)
    if sys.argv[1] == "sample":
        ph = o.threshold(o.read_tiff_raw("/root/reference/data/SampleData_2Phase_stack_3d_1bit.tif"))
        study2(ph, 1, 0, cf)
    else:
        ph = o.sphere_packing(int(sys.argv[1]), radius=int(sys.argv[2])).astype(np.int32)
        study2(ph, 1, 2, cf)

if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "w4":
    vcycle2 = vcycle3
    cf = {}
    cf["d4 lo.15 s.5 V"] = dict(w=cheb_weights(4, .15), cw=cheb_weights(8, 0.05), scale=.5, wdepth=0)
    cf["d3 lo.2 s.5 V"] = dict(w=cheb_weights(3, .2), cw=cheb_weights(8, 0.05), scale=.5, wdepth=0)
    cf["d2 lo.25 s.7 W2"] = dict(w=cheb_weights(2, .25), cw=cheb_weights(8, 0.05), scale=.7, wdepth=2)
    cf["d2 lo.25 s.7 W3"] = dict(w=cheb_weights(2, .25), cw=cheb_weights(8, 0.05), scale=.7, wdepth=3)
    cf["d3 lo.2 s.7 W2"] = dict(w=cheb_weights(3, .2), cw=cheb_weights(8, 0.05), scale=.7, wdepth=2)
    cf["d3 lo.2 s.7 W99"] = dict(w=cheb_weights(3, .2), cw=cheb_weights(8, 0.05), scale=.7, wdepth=99)
    cf["d2 lo.25 s.6 W3"] = dict(w=cheb_weights(2, .25), cw=cheb_weights(8, 0.05), scale=.6, wdepth=3)
    ph = o.sphere_packing(int(sys.argv[1]), radius=int(sys.argv[2])).astype(np.int32)
    study2(ph, 1, 2, cf)
