"""Design study (CPU, scipy): level-0 multigrid vectors STORED in 16 bits (arithmetic in fp32).

The level-0 smoother is HBM bound at 13 B/cell (z 4 + r 4 + flags 1 + z' 4); with z and r held
as fp16 or bf16 it would move 7 B/cell.  The V-cycle is linear, so r is scaled to max|r| = 1
before the cycle and the result scaled back (fp16 has no dynamic range to spare otherwise).
Question: does the rounding of the stored iterates cost PCG iterations?

    python tools/mg_half_study.py 128        # sphere packing n (radius 12), or `sample`
"""
import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import importlib.util
import sys

import numpy as np

sys.path.insert(0, _ROOT)
spec = importlib.util.spec_from_file_location("mgp", os.path.join(_ROOT, "tools/mg_prototype.py"))
m = importlib.util.module_from_spec(spec)
_argv = sys.argv; sys.argv = ["x"]; spec.loader.exec_module(m); sys.argv = _argv
o = m.o


def to_bf16(v):
    """round-to-nearest-even truncation of fp32 to 8 significand bits, returned as fp32"""
    u = np.ascontiguousarray(v, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return (((u + r) >> 16) << 16).astype(np.uint32).view(np.float32)


ROUND = {"fp32": lambda v: v, "fp16": lambda v: v.astype(np.float16).astype(np.float32), "bf16": to_bf16}


def smooth0(L, x, b, weights, rnd):
    A, dinv = L["A"], L["dinv"]
    for w in weights:
        x = w * dinv * b if x is None else x + w * dinv * (b - A @ x)
        x = rnd(x.astype(np.float32))           # the iterate goes back to memory in 16 bits
    return x


def vcycle_half(lev32, r, cfg, rnd):
    L = lev32[0]
    b = rnd(r.astype(np.float32))               # r32 stored in 16 bits
    x = smooth0(L, None, b, cfg["w"], rnd)
    rc = L["P"].T @ (b - L["A"] @ x)
    ec = m.vcycle2(lev32, 1, rc, cfg) if len(lev32) > 1 else 0.0
    x = rnd((x + L["P"] @ ec).astype(np.float32))
    return smooth0(L, x, b, cfg["w"][::-1], rnd)


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
    lev32 = []
    for L in levels:
        d32 = dict(A=L["A"].astype(np.float32), dinv=L["dinv"].astype(np.float32))
        if "P" in L:
            d32["P"] = L["P"].astype(np.float32)
        lev32.append(d32)
    cfg = dict(w=m.cheb_weights(4, .15).astype(np.float32), cw=m.cheb_weights(8, .05).astype(np.float32))
    for name, rnd in ROUND.items():
        def prec(r, rnd=rnd):
            s = np.abs(r).max()
            return vcycle_half(lev32, r / s, cfg, rnd).astype(np.float64) * s
        for eps in (1e-9, 1e-12):
            x, it, h = m.pcg(Auu, bu, x0[unk], prec, eps * bn, maxiter=300)
            xfull = xf.copy(); xfull[unk] = x
            fin, fout, _, _ = o.global_fluxes(xfull.reshape(ph.shape), mask, d)
            tau, _ = o.tau_from_fluxes(fin, fout, mask.sum() / ph.size, ph.shape, d, -1.0, 1.0)
            print(f"{label} level-0 vectors in {name}: eps {eps:g} iters {it:3d} true relres "
                  f"{np.linalg.norm(bu - Auu @ x) / bn:.2e} tau {tau:.10f}", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "96"
if which == "sample":
    ph = o.threshold(o.read_tiff_raw(os.path.join(_ROOT, "tests/golden/SampleData_2Phase_stack_3d_1bit.tif")))
    run(ph, 1, 0, "sample p1 X")
else:
    ph = o.sphere_packing(int(which), radius=12).astype(np.int32)
    run(ph, 1, 2, f"pack{which}")
