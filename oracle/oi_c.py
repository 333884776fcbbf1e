"""ctypes access to oracle/liboi_oracle.so (the plain-C restatement).
TEST INFRASTRUCTURE ONLY -- see the header of oi_oracle.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboi_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "oi_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        lib = C.CDLL(_LIB)
        lib.oo_num_threads.restype = C.c_int
        lib.oo_set_num_threads.restype = C.c_int
        lib.oo_set_num_threads.argtypes = [C.c_int]
        lib.oo_count_phase_i32.restype = C.c_int64
        lib.oo_count_phase_i32.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        lib.oo_remspot.restype = None
        lib.oo_remspot.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.oo_flood_fill.restype = C.c_int
        lib.oo_flood_fill.argtypes = [C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int]
        lib.oo_activity_mask.restype = C.c_int64
        lib.oo_activity_mask.argtypes = [C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_int]
        lib.oo_fillmtx.restype = None
        lib.oo_fillmtx.argtypes = [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p, C.c_double, C.c_double,
                                                                      C.c_int32, C.c_int]
        lib.oo_solve_pcg.restype = C.c_int
        lib.oo_solve_pcg.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_double, C.c_int, C.POINTER(C.c_double)]
        lib.oo_fluxes.restype = None
        lib.oo_fluxes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64)]
        lib.oo_tau.restype = C.c_double
        lib.oo_tau.argtypes = [C.c_double] * 3 + [C.c_int] * 4 + [C.c_void_p, C.c_double, C.c_double, C.c_int,
                                                                  C.POINTER(C.c_double)]
        lib.oo_tortuosity.restype = C.c_int
        lib.oo_tortuosity.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int, C.c_double,
                                      C.c_double, C.c_double, C.c_int, C.c_void_p]
        lib.oo_tortuosity_mg.restype = C.c_int
        lib.oo_tortuosity_mg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int, C.c_double,
                                         C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def num_threads() -> int:
    return load().oo_num_threads()


def set_num_threads(n: int) -> int:
    """OpenMP threads of every later call (overrides OMP_NUM_THREADS); returns the count in force."""
    return load().oo_set_num_threads(int(n))


def host_cores() -> int:
    """Cores this process may run on (the affinity mask, not the machine total)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def count_phase(phase, phase_id):
    p = _i32(phase)
    return int(load().oo_count_phase_i32(p.ctypes.data, p.size, phase_id))


def remspot(phase, passes=1):
    """tortuosity_remspot, in place order; returns the filtered copy."""
    q = _i32(phase).copy()
    nz, ny, nx = q.shape
    for _ in range(passes):
        load().oo_remspot(q.ctypes.data, nx, ny, nz)
    return q


def activity_mask(phase, phase_id, direction, capped=False):
    p = _i32(phase)
    nz, ny, nx = p.shape
    m = np.zeros(p.shape, dtype=np.uint8)
    n = load().oo_activity_mask(p.ctypes.data, phase_id, nx, ny, nz, direction, m.ctypes.data, int(capped))
    return m, int(n)


def flood_fill(phase, phase_id, direction, seed_plane, max_iter=0):
    p = _i32(phase)
    nz, ny, nx = p.shape
    r = np.zeros(p.shape, dtype=np.uint8)
    it = load().oo_flood_fill(p.ctypes.data, phase_id, nx, ny, nz, direction, seed_plane, r.ctypes.data, max_iter)
    return r, it


def fill_matrix(phase, mask, phase_id, direction, vlo, vhi, dx=(1.0, 1.0, 1.0)):
    p = _i32(phase)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    nz, ny, nx = p.shape
    n = p.size
    a = np.empty((n, 7))
    rhs = np.empty(n)
    xinit = np.zeros(n)
    dxinv = np.array([1.0 / d ** 2 for d in dx])
    load().oo_fillmtx(a.ctypes.data, rhs.ctypes.data, xinit.ctypes.data, p.ctypes.data, m.ctypes.data,
                      nx, ny, nz, dxinv.ctypes.data, vlo, vhi, phase_id, direction)
    return a, rhs, xinit


def solve_pcg(a, rhs, x0, shape, eps=1e-9, maxiter=100000):
    nz, ny, nx = shape
    x = np.array(x0, dtype=np.float64).ravel().copy()
    rel = C.c_double(0)
    it = load().oo_solve_pcg(np.ascontiguousarray(a).ctypes.data, np.ascontiguousarray(rhs).ctypes.data,
                             x.ctypes.data, nx, ny, nz, eps, maxiter, C.byref(rel))
    return x.reshape(shape), it, rel.value


def tortuosity(phase, phase_id, direction, vlo=-1.0, vhi=1.0, eps=1e-9, maxiter=100000):
    """-> dict(tau, deff, active_vf, flux_in, flux_out, iters, relres, n_active, solve_s)"""
    p = _i32(phase)
    nz, ny, nx = p.shape
    out = np.zeros(9)
    load().oo_tortuosity(p.ctypes.data, nx, ny, nz, phase_id, direction, vlo, vhi, eps, maxiter, out.ctypes.data)
    keys = ("tau", "deff", "active_vf", "flux_in", "flux_out", "iters", "relres", "n_active", "solve_s")
    d = dict(zip(keys, out.tolist()))
    d["iters"], d["n_active"] = int(d["iters"]), int(d["n_active"])
    return d


def tortuosity_mg(phase, phase_id, direction, vlo=-1.0, vhi=1.0, eps=1e-9, maxiter=200, d0=0, dc=0):
    """The same path with the CPU port of the GPU arm's MG-PCG solver (oo_solve_mgpcg) instead of Jacobi-PCG.
    -> dict(tau, deff, active_vf, flux_in, flux_out, iters, relres, n_active, solve_s, mask_s)"""
    p = _i32(phase)
    nz, ny, nx = p.shape
    out = np.zeros(10)
    load().oo_tortuosity_mg(p.ctypes.data, nx, ny, nz, phase_id, direction, vlo, vhi, eps, maxiter, d0, dc, out.ctypes.data)
    keys = ("tau", "deff", "active_vf", "flux_in", "flux_out", "iters", "relres", "n_active", "solve_s", "mask_s")
    d = dict(zip(keys, out.tolist()))
    d["iters"], d["n_active"] = int(d["iters"]), int(d["n_active"])
    return d


# ---- homogenisation cell problem (second restatement, see oracle/oi_effdiff.py for the numpy one)
def effdiff_fill_matrix(phase, phase_id, direction, dx=(1.0, 1.0, 1.0)):
    lib = load()
    p = _i32(phase)
    nz, ny, nx = p.shape
    n = p.size
    a, rhs, xinit = np.empty((n, 7)), np.empty(n), np.empty(n)
    d = np.array(dx, dtype=np.float64)
    lib.oo_effdiff_fillmtx.restype = None
    lib.oo_effdiff_fillmtx.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.oo_effdiff_fillmtx(a.ctypes.data, rhs.ctypes.data, xinit.ctypes.data, p.ctypes.data, phase_id, nx, ny, nz,
                           d.ctypes.data, direction)
    return a, rhs, xinit


def effdiff_deff_tensor(phase, phase_id, dx=(1.0, 1.0, 1.0), eps=1e-12, maxiter=100000):
    """D_eff / D by the C restatement: three Jacobi-PCG corrector solves + gradient sums."""
    lib = load()
    p = _i32(phase)
    nz, ny, nx = p.shape
    d = np.array(dx, dtype=np.float64)
    lib.oo_effdiff_solve.restype = C.c_int
    lib.oo_effdiff_solve.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_double, C.c_int, C.POINTER(C.c_double)]
    lib.oo_effdiff_gradient_sums.restype = None
    lib.oo_effdiff_gradient_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.POINTER(C.c_int64)]
    D = np.zeros((3, 3))
    iters = []
    for k in range(3):
        a, rhs, x = effdiff_fill_matrix(p, phase_id, k, dx)
        rel = C.c_double(0)
        iters.append(lib.oo_effdiff_solve(a.ctypes.data, rhs.ctypes.data, x.ctypes.data, nx, ny, nz, eps, maxiter,
                                          C.byref(rel)))
        sums = np.zeros(3)
        na = C.c_int64(0)
        lib.oo_effdiff_gradient_sums(x.ctypes.data, p.ctypes.data, phase_id, nx, ny, nz, d.ctypes.data,
                                     sums.ctypes.data, C.byref(na))
        for ax in range(3):
            D[ax][k] = ((na.value if ax == k else 0.0) - sums[ax]) / p.size
    return D, iters
