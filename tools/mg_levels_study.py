"""Scratch study: smoother degree per level (fine vs coarse) on a sphere packing."""
import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys, time
import numpy as np
sys.path.insert(0, _ROOT)
sys.argv = [sys.argv[0]] + sys.argv[1:]
import importlib.util
spec = importlib.util.spec_from_file_location("mgp", os.path.join(_ROOT, "tools/mg_prototype.py"))
m = importlib.util.module_from_spec(spec)
_argv = sys.argv; sys.argv = ["x"]; spec.loader.exec_module(m); sys.argv = _argv
o = m.o

LO = {1: .4, 2: .25, 3: .2, 4: .15, 5: .12, 6: .1}

def vcyc(levels, l, b, degs, cw):
    L = levels[l]
    if l == len(levels) - 1:
        return m.smooth_wjac(L, None, b, cw)
    w = m.cheb_weights(degs[min(l, len(degs) - 1)], LO[degs[min(l, len(degs) - 1)]])
    x = m.smooth_wjac(L, None, b, w)
    r = b - L["A"] @ x
    ec = vcyc(levels, l + 1, L["P"].T @ r, degs, cw)
    x = x + L["P"] @ ec
    return m.smooth_wjac(L, x, b, w[::-1])

n = int(sys.argv[1]); R = int(sys.argv[2])
ph = o.sphere_packing(n, radius=R).astype(np.int32)
mask = o.activity_mask(ph, 1, 2)
a, rhs, x0 = o.fill_matrix(ph, mask, 1, 2, -1.0, 1.0)
A = o.assemble_csr(a, ph.shape)
Auu, bu, unk, xf = o.eliminate_dirichlet(A, rhs, x0, ph.shape, mask, 2)
bn = o.reference_stop_norm(rhs)
nz, ny, nx = ph.shape
lin = np.nonzero(unk)[0]
idx3 = (lin // (nx * ny), (lin // nx) % ny, lin % nx)
levels = m.build_hierarchy(None, Auu, idx3, ph.shape, scale=0.5)
cw = m.cheb_weights(8, 0.05)
for degs in ([4], [4, 3], [4, 2], [4, 3, 2], [3, 3], [5, 3], [4, 4, 2], [6, 2], [5, 2]):
    t = time.time()
    x, it, h = m.pcg(Auu, bu, x0[unk], lambda r: vcyc(levels, 0, r, degs, cw), 1e-9 * bn, maxiter=300)
    print(degs, "iters", it, f"{time.time()-t:.0f}s", flush=True)
