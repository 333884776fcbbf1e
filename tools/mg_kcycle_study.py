"""Design study (CPU, scipy): would Krylov-accelerated coarse solves (the K-cycle of Notay /
Vassilevski, as in AGMG) pay on top of the shipped V-cycle?  Plain 2x2x2 aggregation with
piecewise-constant transfers is not h-independent (21 / 26 / 33 PCG iterations at 512^3 /
1024^3 / 2048^3 on the GPU); a K-cycle solves every coarse problem with two steps of
flexible CG preconditioned by the next level's cycle, which is known to restore a
level-independent rate for aggregation hierarchies at W-cycle-like cost.

Prints, per configuration, the outer (flexible) PCG iteration count to the reference's
stopping rule and the visits per level, from which DESIGN.md's cost estimate is made.

    python tools/mg_kcycle_study.py 128 12        # sphere packing n, radius
"""
import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys
import time

import numpy as np

sys.path.insert(0, _ROOT)
sys.argv, _argv = sys.argv[:1], sys.argv          # keep mg_prototype's __main__ blocks quiet
from tools.mg_prototype import build_hierarchy, cheb_weights, smooth_wjac  # noqa: E402
from oracle import oi_numpy as o  # noqa: E402

sys.argv = _argv
VISITS = {}


def cycle(levels, l, b, cfg):
    """One multigrid cycle at level l (zero initial guess)."""
    VISITS[l] = VISITS.get(l, 0) + 1
    L = levels[l]
    if l == len(levels) - 1:
        return smooth_wjac(L, None, b, cfg["cw"])
    x = smooth_wjac(L, None, b, cfg["w"])
    r = b - L["A"] @ x
    rc = L["P"].T @ r
    if l + 1 >= cfg.get("kdepth", 99) and l + 1 < len(levels) - 1:
        ec = kcycle_solve(levels, l + 1, rc, cfg)
    else:
        ec = cycle(levels, l + 1, rc, cfg)
    x = x + L["P"] @ ec
    return smooth_wjac(L, x, b, cfg["w"][::-1])


def kcycle_solve(levels, l, r, cfg):
    """Two steps of flexible CG on A_l e = r, preconditioned by cycle(l) (Notay & Vassilevski 2008)."""
    A = levels[l]["A"]
    c1 = cycle(levels, l, r, cfg)
    v1 = A @ c1
    rho1, alpha1 = c1 @ v1, c1 @ r
    if rho1 <= 0:
        return c1
    r1 = r - (alpha1 / rho1) * v1
    if np.linalg.norm(r1) <= cfg.get("kt", 0.25) * np.linalg.norm(r):
        return (alpha1 / rho1) * c1
    c2 = cycle(levels, l, r1, cfg)
    v2 = A @ c2
    gamma, beta, alpha2 = c2 @ v1, c2 @ v2, c2 @ r1
    rho2 = beta - gamma * gamma / rho1
    if rho2 <= 0:
        return (alpha1 / rho1) * c1
    return (alpha1 / rho1 - gamma * alpha2 / (rho1 * rho2)) * c1 + (alpha2 / rho2) * c2


def fpcg(A, b, x0, prec, tol_abs, maxiter=300):
    """Flexible PCG (Polak-Ribiere beta): the K-cycle is a non-linear preconditioner."""
    x = x0.copy()
    r = b - A @ x
    z = prec(r)
    p = z.copy()
    rz = r @ z
    for it in range(1, maxiter + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r_old = r.copy()
        r -= alpha * Ap
        if np.linalg.norm(r) <= tol_abs:
            return x, it
        z = prec(r)
        rz_new = r @ z
        beta = (z @ (r - r_old)) / rz
        p = z + beta * p
        rz = rz_new
    return x, maxiter


def study(phase, phase_id, direction, cfgs, eps=1e-9):
    mask = o.activity_mask(phase, phase_id, direction)
    a, rhs, x0 = o.fill_matrix(phase, mask, phase_id, direction, -1.0, 1.0)
    A = o.assemble_csr(a, phase.shape)
    Auu, bu, unk, xf = o.eliminate_dirichlet(A, rhs, x0, phase.shape, mask, direction)
    bnorm = o.reference_stop_norm(rhs)
    nz, ny, nx = phase.shape
    lin = np.nonzero(unk)[0]
    idx3 = (lin // (nx * ny), (lin // nx) % ny, lin % nx)
    print(f"shape {phase.shape} unknowns {Auu.shape[0]}", flush=True)
    cache = {}
    for name, cfg in cfgs.items():
        t = time.time()
        sc = cfg.get("scale", 0.5)
        if sc not in cache:
            cache[sc] = build_hierarchy(None, Auu, idx3, phase.shape, scale=sc)
        levels = cache[sc]
        VISITS.clear()
        x, it = fpcg(Auu, bu, x0[unk], lambda r: cycle(levels, 0, r, cfg), eps * bnorm)
        per_it = {l: round(v / (it + 1), 2) for l, v in sorted(VISITS.items())}
        print(f"  {name:30s} iters {it:3d}  visits/iter by level {per_it}  {time.time() - t:.0f}s", flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rad = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    ph = o.sphere_packing(n, radius=rad).astype(np.int32)
    w4, cw = cheb_weights(4, 0.15), cheb_weights(8, 0.05)
    w2, w3 = cheb_weights(2, 0.25), cheb_weights(3, 0.2)
    if len(sys.argv) > 3 and sys.argv[3] == "short":
        cfgs = {
            "V d4 s.5 (shipped)": dict(w=w4, cw=cw, scale=0.5),
            "K from level 1, d4 s1": dict(w=w4, cw=cw, scale=1.0, kdepth=1),
            "K from level 1, d4 s.5 always 2": dict(w=w4, cw=cw, scale=0.5, kdepth=1, kt=0.0),
            "K from level 1, d4 s.7": dict(w=w4, cw=cw, scale=0.7, kdepth=1),
            "K from level 1, d3 s1": dict(w=w3, cw=cw, scale=1.0, kdepth=1),
        }
        study(ph, 1, 2, cfgs)
        sys.exit(0)
    cfgs = {
        "V d4 s.5 (shipped)": dict(w=w4, cw=cw, scale=0.5),
        "K from level 1, d4 s.5": dict(w=w4, cw=cw, scale=0.5, kdepth=1),
        "K from level 1, d4 s1": dict(w=w4, cw=cw, scale=1.0, kdepth=1),
        "K from level 2, d4 s.5": dict(w=w4, cw=cw, scale=0.5, kdepth=2),
        "K from level 2, d4 s1": dict(w=w4, cw=cw, scale=1.0, kdepth=2),
        "K from level 1, d2 s.5": dict(w=w2, cw=cw, scale=0.5, kdepth=1),
        "K from level 1, d3 s.5": dict(w=w3, cw=cw, scale=0.5, kdepth=1),
        "K from level 1, d2 s1": dict(w=w2, cw=cw, scale=1.0, kdepth=1),
    }
    study(ph, 1, 2, cfgs)
