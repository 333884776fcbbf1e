"""Scratch study: V-cycle carried out in fp32 (operators + vectors) inside an fp64 PCG."""
import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys, time
import numpy as np
sys.path.insert(0, _ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("mgp", os.path.join(_ROOT, "tools/mg_prototype.py"))
m = importlib.util.module_from_spec(spec)
_argv = sys.argv; sys.argv = ["x"]; spec.loader.exec_module(m); sys.argv = _argv
o = m.o

def run(ph, pid, d, label):
    mask = o.activity_mask(ph, pid, d)
    a, rhs, x0 = o.fill_matrix(ph, mask, pid, d, -1.0, 1.0)
    A = o.assemble_csr(a, ph.shape)
    Auu, bu, unk, xf = o.eliminate_dirichlet(A, rhs, x0, ph.shape, mask, d)
    bn = o.reference_stop_norm(rhs)
    nz, ny, nx = ph.shape
    lin = np.nonzero(unk)[0]
    idx3 = (lin // (nx * ny), (lin // nx) % ny, lin % nx)
    levels = m.build_hierarchy(None, Auu, idx3, ph.shape, scale=0.5)
    cfg = dict(w=m.cheb_weights(4, .15), cw=m.cheb_weights(8, .05))
    lev32 = []
    for L in levels:
        d32 = dict(A=L["A"].astype(np.float32), dinv=L["dinv"].astype(np.float32))
        if "P" in L: d32["P"] = L["P"].astype(np.float32)
        lev32.append(d32)
    cfg32 = dict(w=cfg["w"].astype(np.float32), cw=cfg["cw"].astype(np.float32))
    for name, prec in (("fp64", lambda r: m.vcycle2(levels, 0, r, cfg)),
                       ("fp32", lambda r: m.vcycle2(lev32, 0, r.astype(np.float32), cfg32).astype(np.float64)),
                       ("fp32 scaled", None)):
        if prec is None:
            def prec(r):
                s = np.abs(r).max()
                return m.vcycle2(lev32, 0, (r / s).astype(np.float32), cfg32).astype(np.float64) * s
        for eps in (1e-9, 1e-12):
            x, it, h = m.pcg(Auu, bu, x0[unk], prec, eps * bn, maxiter=300)
            xfull = xf.copy(); xfull[unk] = x
            fin, fout, _, _ = o.global_fluxes(xfull.reshape(ph.shape), mask, d)
            tau, _ = o.tau_from_fluxes(fin, fout, mask.sum() / ph.size, ph.shape, d, -1.0, 1.0)
            print(f"{label} {name:12s} eps {eps:g}: iters {it} true relres {np.linalg.norm(bu - Auu @ x)/bn:.2e} tau {tau:.10f}", flush=True)

which = sys.argv[1]
if which == "sample":
    ph = o.threshold(o.read_tiff_raw(os.path.join(_ROOT, "tests/golden/SampleData_2Phase_stack_3d_1bit.tif")))
    run(ph, 1, 0, "sample p1 X")
else:
    ph = o.sphere_packing(int(which), radius=12).astype(np.int32)
    run(ph, 1, 2, f"pack{which}")
