"""numpy/scipy CPU restatement of the reference homogenisation path
(EffectiveDiffusivityHypre + effdiff_fillmtx + calculate_Deff_tensor_homogenization).

TEST INFRASTRUCTURE ONLY (see oracle/oi_numpy.py): imported by tests/ as the
checker of the CUDA cell-problem path, never by the product.

Parity status: "parity unpinned" against real HYPRE output -- the reference
ships no test or golden value for this path (tests/ has no tEffectiveDiffusivity
numbers) and cannot be built here.  What is pinned: the assembled rows follow
src/props/EffDiffFillMtx.F90 statement by statement (checked bit-for-bit against
the rows the CUDA path exports), the system is solved to 1e-12 by an independent
Krylov solver (scipy CG on the assembled CSR), and the analytic cases: an
all-pore periodic box has chi = 0 and D_eff = I; a box of pore/solid layers
normal to x has D_xx = 0 ... see tests/test_gpu_effdiff.py.

Arrays are indexed [k, j, i] = [z, y, x], x fastest.  Citations are file:line
relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

# stencil slot order C,-x,+x,-y,+y,-z,+z (src/props/EffDiffFillMtx.F90:31-37)
_AXIS_OF_SLOT = (None, 2, 2, 1, 1, 0, 0)       # numpy axis of the neighbour
_SHIFT_OF_SLOT = (0, 1, -1, 1, -1, 1, -1)      # np.roll shift that brings the neighbour here


def active_mask(phase: np.ndarray, phase_id: int) -> np.ndarray:
    """generateActiveMask (src/props/EffectiveDiffusivityHypre.cpp:213-330):
    phase == id, no percolation filter; ghosts are filled periodically."""
    return (phase == phase_id)


def fill_matrix(phase: np.ndarray, phase_id: int, direction: int, dx=(1.0, 1.0, 1.0)):
    """effdiff_fillmtx (src/props/EffDiffFillMtx.F90:109-258) on the whole periodic
    box.  Returns (a[N,7], rhs[N], xinit[N]), cell index x-fastest."""
    act = active_mask(phase, phase_id)
    nz, ny, nx = phase.shape
    inv_d2 = (1.0 / (dx[0] * dx[0]), 1.0 / (dx[1] * dx[1]), 1.0 / (dx[2] * dx[2]))   # F90:93-95
    coef = (0.0, inv_d2[0], inv_d2[0], inv_d2[1], inv_d2[1], inv_d2[2], inv_d2[2])
    a = np.zeros((nz, ny, nx, 7))
    diag = np.zeros((nz, ny, nx))
    flux = np.zeros((nz, ny, nx))
    nb_act = [None] * 7
    for s in range(1, 7):
        nb = np.roll(act, _SHIFT_OF_SLOT[s], axis=_AXIS_OF_SLOT[s])     # periodic FillBoundary
        nb_act[s] = nb
        a[..., s] = np.where(act & nb, -coef[s], 0.0)                   # F90:152-154 etc.
        diag += coef[s]                                                 # both branches add inv_dx2 (F90:153,156)
    # interface faces along `direction` feed the rhs (F90:157-159, 167-169, ...)
    h = dx[direction]
    s_m, s_p = 1 + 2 * direction, 2 + 2 * direction
    flux = flux + np.where(~nb_act[s_m], 1.0 / h, 0.0)
    flux = flux - np.where(~nb_act[s_p], 1.0 / h, 0.0)
    inv_2h = 1.0 / (2.0 * h)                                            # F90:98-100
    div = -(nb_act[s_p].astype(np.float64) - nb_act[s_m].astype(np.float64)) * inv_2h   # F90:226-234
    rhs = np.where(act, div + flux, 0.0)
    a[..., 0] = diag
    inact = ~act                                                        # F90:124-129
    a[inact] = 0.0
    a[inact, 0] = 1.0
    N = nx * ny * nz
    return a.reshape(N, 7), rhs.reshape(N), np.zeros(N)


def assemble_csr_periodic(a: np.ndarray, shape):
    """CSR of the periodic 7-point struct matrix (HYPRE_StructGridSetPeriodic,
    src/props/EffectiveDiffusivityHypre.cpp:341-368)."""
    import scipy.sparse as sp
    nz, ny, nx = shape
    N = nx * ny * nz
    m = np.arange(N)
    i, j, k = m % nx, (m // nx) % ny, m // (nx * ny)
    nbr = (m,
           k * ny * nx + j * nx + (i - 1) % nx, k * ny * nx + j * nx + (i + 1) % nx,
           k * ny * nx + ((j - 1) % ny) * nx + i, k * ny * nx + ((j + 1) % ny) * nx + i,
           ((k - 1) % nz) * ny * nx + j * nx + i, ((k + 1) % nz) * ny * nx + j * nx + i)
    rows, cols, vals = [], [], []
    for s in range(7):
        v = a[:, s]
        sel = v != 0.0
        rows.append(m[sel]); cols.append(nbr[s][sel]); vals.append(v[sel])
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    return A.tocsr()            # duplicates (n = 1 or 2 along an axis) are summed


def solve_chi(phase: np.ndarray, phase_id: int, direction: int, dx=(1.0, 1.0, 1.0), eps: float = 1e-12):
    """chi_k on the whole box (zero in the solid).  Returns (chi[k,j,i], iterations, relres)."""
    import scipy.sparse.linalg as spla
    a, rhs, x0 = fill_matrix(phase, phase_id, direction, dx)
    A = assemble_csr_periodic(a, phase.shape)
    bn = float(np.linalg.norm(rhs))
    if bn == 0.0:
        return np.zeros(phase.shape), 0, 0.0
    its = [0]

    def cb(_):
        its[0] += 1
    x, info = spla.cg(A, rhs, x0=x0, rtol=eps, atol=0.0, maxiter=20000, callback=cb)
    rel = float(np.linalg.norm(rhs - A @ x)) / bn
    return x.reshape(phase.shape), its[0], rel


def gradient_sums(chi: np.ndarray, act: np.ndarray, dx=(1.0, 1.0, 1.0)):
    """sum over active cells of the central differences of chi (periodic), the
    per-direction ingredient of calculate_Deff_tensor_homogenization
    (src/props/Diffusion.cpp:115-131)."""
    out = []
    for axis_xyz in range(3):
        ax = 2 - axis_xyz
        g = (np.roll(chi, -1, axis=ax) - np.roll(chi, 1, axis=ax)) * (1.0 / (2.0 * dx[axis_xyz]))
        out.append(float(g[act].sum()))
    return tuple(out)


def deff_tensor(phase: np.ndarray, phase_id: int, dx=(1.0, 1.0, 1.0), eps: float = 1e-12):
    """D_eff / D (src/props/Diffusion.cpp:60-167): D[a][k] = sum_active (delta_ak - d chi_k / d x_a) / N."""
    act = active_mask(phase, phase_id)
    n_act = int(act.sum())
    D = np.zeros((3, 3))
    for k in range(3):
        chi, _, _ = solve_chi(phase, phase_id, k, dx, eps)
        s = gradient_sums(chi, act, dx)
        for a_ in range(3):
            D[a_][k] = ((n_act if a_ == k else 0.0) - s[a_]) / phase.size
    return D
