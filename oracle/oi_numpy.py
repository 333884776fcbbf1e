"""numpy/scipy CPU restatement of the reference TortuosityHypre path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import it, and only as the checker.

Parity status: the reference's own tests pin *no* numeric VF/tau values
(tests/inputs/tTortuosity.inputs:30-32 are commented out) and the reference
(AMReX + HYPRE + MPI + gfortran) cannot be built in this container, so the
fp64 solve is "parity unpinned" against real HYPRE output.  What IS pinned:
the reader facts of src/io/tTiffReader.cpp:98-219 (100^3, 1 bit), the
checkMatrixProperties invariants (src/props/TortuosityHypre.cpp:896-982) on
the assembled rows, and analytic known-answer cases (tau = (N-1)/N).

All arrays are indexed [k, j, i] = [z, y, x] with x fastest in memory, which
is the AMReX/Fortran (i fastest) layout used by the reference.
Citations are file:line relative to /root/reference.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

# --------------------------------------------------------------------------
# a-1  reader: TIFF decode + threshold  (src/io/TiffReader.cpp:289-444)
# --------------------------------------------------------------------------

_TIFF_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h",
               9: "i", 10: "ii", 11: "f", 12: "d", 16: "Q", 17: "q", 18: "Q"}


def _read_ifds(buf: bytes):
    """Walk every IFD of a classic or BigTIFF file -> list of {tag: values}."""
    bo = {b"II": "<", b"MM": ">"}[buf[:2]]
    magic = struct.unpack(bo + "H", buf[2:4])[0]
    big = magic == 43
    if big:
        off = struct.unpack(bo + "Q", buf[8:16])[0]
    else:
        off = struct.unpack(bo + "I", buf[4:8])[0]
    ifds = []
    while off:
        if big:
            n = struct.unpack(bo + "Q", buf[off:off + 8])[0]
            p, esz, cfmt, vsz = off + 8, 20, "Q", 8
        else:
            n = struct.unpack(bo + "H", buf[off:off + 2])[0]
            p, esz, cfmt, vsz = off + 2, 12, "I", 4
        tags = {}
        for e in range(n):
            ent = buf[p + e * esz:p + (e + 1) * esz]
            tag, typ = struct.unpack(bo + "HH", ent[:4])
            cnt = struct.unpack(bo + cfmt, ent[4:4 + vsz])[0]
            f = _TIFF_TYPES.get(typ)
            if f is None:
                continue
            per = struct.calcsize("=" + f)
            nbytes = per * cnt
            if nbytes <= vsz:
                raw = ent[4 + vsz:4 + vsz + nbytes]
            else:
                o = struct.unpack(bo + cfmt, ent[4 + vsz:4 + 2 * vsz])[0]
                raw = buf[o:o + nbytes]
            if f == "c":
                tags[tag] = raw
            else:
                tags[tag] = struct.unpack(bo + f * cnt, raw)
        ifds.append(tags)
        q = p + n * esz
        off = struct.unpack(bo + cfmt, buf[q:q + vsz])[0]
    return bo, ifds


def read_tiff_raw(path: str) -> np.ndarray:
    """Raw sample values of an uncompressed multi-directory TIFF as float64
    [z, y, x] exactly as the reference decodes them: 1-bit samples unpacked
    with the TIFFScanlineSize row pitch ceil(W/8), MSB first when
    FillOrder==1 (TiffReader.cpp:419-426), wider samples through
    interpretBytesAsDouble (TiffReader.cpp:41-79); PhotometricInterpretation
    is NOT applied (no inversion anywhere in TiffReader.cpp)."""
    buf = open(path, "rb").read()
    bo, ifds = _read_ifds(buf)
    t0 = ifds[0]
    W, H = t0[256][0], t0[257][0]
    bps = t0.get(258, (1,))[0]
    fmt = t0.get(339, (1,))[0]
    spp = t0.get(277, (1,))[0]
    if spp != 1 or t0.get(284, (1,))[0] != 1:
        raise ValueError("unsupported TIFF (SPP/planar)")  # TiffReader.cpp:168
    if bps not in (1, 8, 16, 32, 64):
        raise ValueError("unsupported BitsPerSample")      # TiffReader.cpp:167
    out = np.zeros((len(ifds), H, W), dtype=np.float64)
    for k, t in enumerate(ifds):
        if t.get(259, (1,))[0] != 1:
            raise ValueError("compressed TIFF not supported by the oracle")
        fill = t.get(266, (1,))[0]
        if 322 in t:  # tiled (TiffReader.cpp:354-393)
            tw, th = t[322][0], t[323][0]
            offs, cnts = t[324], t[325]
            tiles_x = (W + tw - 1) // tw
            for ti, (o, c) in enumerate(zip(offs, cnts)):
                data = np.frombuffer(buf[o:o + c], dtype=np.uint8)
                ty, tx = divmod(ti, tiles_x)
                y0, x0 = ty * th, tx * tw
                hh, ww = min(th, H - y0), min(tw, W - x0)
                if bps == 1:
                    # libtiff hands back MSB2LSB data; a FillOrder=2 file is
                    # bit-reversed by libtiff and then read LSB-first by the
                    # reference (TiffReader.cpp:380) == MSB-first raw bits.
                    # NB the reference indexes bits of a tile linearly with
                    # pitch tile_width (TiffReader.cpp:378), not byte-padded.
                    bits = np.unpackbits(data)
                    lin = (np.arange(hh)[:, None] * tw + np.arange(ww)[None, :])
                    ok = (lin // 8) < len(data)
                    vals = np.where(ok, bits[np.minimum(lin, len(bits) - 1)], 0)
                else:
                    vals = _samples(data, bps, fmt, bo)[:th * tw].reshape(th, tw)[:hh, :ww]
                out[k, y0:y0 + hh, x0:x0 + ww] = vals
        else:         # strips (TiffReader.cpp:394-437)
            rps = t.get(278, (H,))[0]
            if rps == 0 or rps > H:
                rps = H
            offs, cnts = t[273], t[279]
            for s, (o, c) in enumerate(zip(offs, cnts)):
                y0 = s * rps
                rows = min(rps, H - y0)
                data = np.frombuffer(buf[o:o + c], dtype=np.uint8)
                if bps == 1:
                    pitch = (W + 7) // 8   # TIFFScanlineSize
                    rowsb = data[:rows * pitch].reshape(rows, pitch)
                    vals = np.unpackbits(rowsb, axis=1)[:, :W]
                else:
                    vals = _samples(data, bps, fmt, bo)[:rows * W].reshape(rows, W)
                out[k, y0:y0 + rows, :] = vals
    return out


def _samples(data: np.ndarray, bps: int, fmt: int, bo: str) -> np.ndarray:
    # libtiff byte-swaps samples to host order for 16/32/64-bit data, so the
    # reference's memcpy (TiffReader.cpp:41-79) sees native values.
    kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
    if kind is None:
        return np.zeros(len(data) // (bps // 8))
    if kind == "f" and bps not in (32, 64):
        return np.zeros(len(data) // (bps // 8))
    n = (len(data) // (bps // 8)) * (bps // 8)
    return data[:n].view(np.dtype(f"{bo}{kind}{bps // 8}")).astype(np.float64)


def threshold(raw: np.ndarray, thr: float = 0.5, v_true: int = 1, v_false: int = 0) -> np.ndarray:
    """phase = (double(raw) > thr) ? v_true : v_false  (TiffReader.cpp:434,
    RawReader.cpp / HDF5Reader.cpp use the same rule)."""
    return np.where(raw.astype(np.float64) > thr, v_true, v_false).astype(np.int32)


def read_raw_file(path: str, nx: int, ny: int, nz: int, dtype="u1") -> np.ndarray:
    """Flat binary, x fastest, z slowest (src/io/RawReader.cpp:310-313)."""
    a = np.fromfile(path, dtype=np.dtype(dtype), count=nx * ny * nz)
    return a.reshape(nz, ny, nx).astype(np.float64)


# --------------------------------------------------------------------------
# a-2  VolumeFraction::value  (src/props/VolumeFraction.cpp:22-66)
# --------------------------------------------------------------------------

def volume_fraction_counts(phase: np.ndarray, phase_id: int):
    """(phase_count, total_count) as exact Python ints."""
    return int(np.count_nonzero(phase == phase_id)), int(phase.size)


# --------------------------------------------------------------------------
# a-3  tortuosity_remspot  (src/props/Tortuosity_filcc.F90:88-177)
# --------------------------------------------------------------------------

def remspot(phase: np.ndarray, passes: int = 1) -> np.ndarray:
    """Isolated-voxel flip, in place, i fastest / k slowest sweep order (the
    Fortran loop order makes it Gauss-Seidel: later cells see earlier flips).
    Small inputs only (pure Python loop)."""
    q = phase.copy()
    nz, ny, nx = q.shape
    for _ in range(passes):
        for k in range(nz):
            for j in range(ny):
                for i in range(nx):
                    v = q[k, j, i]
                    same = False
                    for dk, dj, di in ((0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)):
                        kk, jj, ii = k + dk, j + dj, i + di
                        if 0 <= kk < nz and 0 <= jj < ny and 0 <= ii < nx and q[kk, jj, ii] == v:
                            same = True
                            break
                    if not same:
                        q[k, j, i] = 1 - v if v in (0, 1) else v
    return q


# --------------------------------------------------------------------------
# a-4  generateActivityMask / parallelFloodFill
#      (src/props/TortuosityHypre.cpp:394-558, 297-389)
# --------------------------------------------------------------------------

def _plane(a: np.ndarray, d: int, idx: int):
    """View of the plane idx_dir == idx; d: 0=x,1=y,2=z (axis 2-d)."""
    sl = [slice(None)] * 3
    sl[2 - d] = idx
    return a[tuple(sl)]


def flood_fill(is_phase: np.ndarray, seeds: np.ndarray) -> np.ndarray:
    """Exact fixed point of the reference's sweep (TortuosityHypre.cpp:336-380):
    every phase cell 6-connected to a seed.  Jacobi-style whole-array
    dilation; returns (reached, n_sweeps)."""
    reached = seeds & is_phase
    sweeps = 0
    while True:
        grow = reached.copy()
        grow[1:, :, :] |= reached[:-1, :, :]
        grow[:-1, :, :] |= reached[1:, :, :]
        grow[:, 1:, :] |= reached[:, :-1, :]
        grow[:, :-1, :] |= reached[:, 1:, :]
        grow[:, :, 1:] |= reached[:, :, :-1]
        grow[:, :, :-1] |= reached[:, :, 1:]
        grow &= is_phase
        sweeps += 1
        if np.array_equal(grow, reached):
            return reached, sweeps
        reached = grow


def activity_mask(phase: np.ndarray, phase_id: int, direction: int) -> np.ndarray:
    """mask = reachedInlet AND reachedOutlet (TortuosityHypre.cpp:526-538);
    empty when either face has no seed (TortuosityHypre.cpp:508-514).
    Uses scipy.ndimage.label (6-connectivity) == the flood fixed point."""
    from scipy import ndimage
    is_phase = phase == phase_id
    n = phase.shape[2 - direction]
    lab, _ = ndimage.label(is_phase)  # default structure = 6-connected in 3-D
    lo = np.unique(_plane(lab, direction, 0))
    hi = np.unique(_plane(lab, direction, n - 1))
    lo, hi = lo[lo > 0], hi[hi > 0]
    if len(lo) == 0 or len(hi) == 0:
        return np.zeros_like(is_phase)
    both = np.intersect1d(lo, hi)
    return np.isin(lab, both) & is_phase


def activity_mask_flood(phase: np.ndarray, phase_id: int, direction: int):
    """Same mask via two literal flood fills (slow; cross-check only)."""
    is_phase = phase == phase_id
    n = phase.shape[2 - direction]
    s_in = np.zeros_like(is_phase)
    _plane(s_in, direction, 0)[...] = True
    s_out = np.zeros_like(is_phase)
    _plane(s_out, direction, n - 1)[...] = True
    if not (s_in & is_phase).any() or not (s_out & is_phase).any():
        return np.zeros_like(is_phase), 0
    r_in, n1 = flood_fill(is_phase, s_in)
    r_out, n2 = flood_fill(is_phase, s_out)
    return r_in & r_out, max(n1, n2)


# --------------------------------------------------------------------------
# a-6  tortuosity_fillmtx  (src/props/TortuosityHypreFill.F90:44-314)
# --------------------------------------------------------------------------

# stencil slot order C,-x,+x,-y,+y,-z,+z (TortuosityHypreFill.F90:20-26)
_OFFS = ((0, 0, 0), (0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0))


def fill_matrix(phase: np.ndarray, mask: np.ndarray, phase_id: int, direction: int,
                vlo: float, vhi: float, dx=(1.0, 1.0, 1.0)):
    """Vectorised restatement of tortuosity_fillmtx on the whole domain as one
    box.  Returns (a[N,7], rhs[N], xinit[N]) with cell index m x-fastest.
    The out-of-domain ghost layer is inactive (mask ghosts stay 0:
    TortuosityHypre.cpp:309,522 + non-periodic FillBoundary)."""
    nz, ny, nx = phase.shape
    act = (phase == phase_id) & mask.astype(bool)             # F90:111
    pad = np.zeros((nz + 2, ny + 2, nx + 2), dtype=bool)
    pad[1:-1, 1:-1, 1:-1] = act
    c = (1.0 / dx[0] ** 2, 1.0 / dx[1] ** 2, 1.0 / dx[2] ** 2)  # TortuosityHypre.cpp:580-582
    coef = (0.0, c[0], c[0], c[1], c[1], c[2], c[2])
    a = np.zeros((nz, ny, nx, 7))
    diag = np.zeros((nz, ny, nx))
    for s in range(1, 7):
        dk, dj, di = _OFFS[s]
        nb = pad[1 + dk:1 + dk + nz, 1 + dj:1 + dj + ny, 1 + di:1 + di + nx]
        on = act & nb                                          # F90:126-166
        a[..., s] = np.where(on, -coef[s], 0.0)
        diag += np.where(on, coef[s], 0.0)
    a[..., 0] = diag
    rhs = np.zeros((nz, ny, nx))
    xinit = np.zeros((nz, ny, nx))
    # inactive rows: identity, b = 0, x0 = 0 (F90:111-118)
    inact = ~act
    a[inact] = 0.0
    a[inact, 0] = 1.0
    # active with ~zero diagonal: decoupled identity, x0 = 0 (F90:172-181)
    zero_diag = act & (np.abs(diag) < 1e-15)
    a[zero_diag] = 0.0
    a[zero_diag, 0] = 1.0
    # Dirichlet overwrite on the flow-direction faces (F90:192-228)
    n = phase.shape[2 - direction]
    idx = np.arange(n).reshape([-1 if ax == 2 - direction else 1 for ax in range(3)])
    idx = np.broadcast_to(idx, phase.shape)
    live = act & ~zero_diag
    dlo = live & (idx == 0)
    dhi = live & (idx == n - 1) & ~dlo
    for sel, v in ((dlo, vlo), (dhi, vhi)):
        a[sel] = 0.0
        a[sel, 0] = 1.0
        rhs[sel] = v
    on_dir = dlo | dhi
    # initial guess: linear ramp on live cells whose diagonal != 1, or that are
    # Dirichlet (F90:233-262); the caller's buffer is value-initialised so
    # skipped cells stay 0 (std::vector::resize, TortuosityHypre.cpp:606).
    ramp_sel = live & ((np.abs(a[..., 0] - 1.0) > 1e-15) | on_dir)
    ext = float(n - 1)
    factor = 0.0 if abs(ext) < 1e-15 else 1.0 / ext
    ramp = vlo + (vhi - vlo) * idx.astype(np.float64) * factor
    xinit[ramp_sel] = ramp[ramp_sel]
    N = nx * ny * nz
    return a.reshape(N, 7), rhs.reshape(N), xinit.reshape(N)


def check_matrix_properties(a, rhs, mask, direction, vlo, vhi, shape, tol=1e-14) -> bool:
    """The reference's own invariants (TortuosityHypre.cpp:896-982)."""
    nz, ny, nx = shape
    a = a.reshape(nz, ny, nx, 7)
    rhs = rhs.reshape(nz, ny, nx)
    m = mask.astype(bool)
    if not (np.isfinite(a).all() and np.isfinite(rhs).all()):
        return False
    n = shape[2 - direction]
    idx = np.arange(n).reshape([-1 if ax == 2 - direction else 1 for ax in range(3)])
    idx = np.broadcast_to(idx, shape)
    dirichlet = m & ((idx == 0) | (idx == n - 1))
    inactive = ~m
    interior = m & ~dirichlet
    ok = True
    ok &= bool((np.abs(a[inactive, 0] - 1.0) <= tol).all() and (np.abs(rhs[inactive]) <= tol).all())
    ok &= bool((np.abs(a[inactive, 1:]) <= tol).all())
    exp = np.where(idx == 0, vlo, vhi)
    ok &= bool((np.abs(a[dirichlet, 0] - 1.0) <= tol).all() and (np.abs(rhs[dirichlet] - exp[dirichlet]) <= tol).all())
    ok &= bool((np.abs(a[dirichlet, 1:]) <= tol).all())
    ok &= bool((a[interior, 0] > tol).all() and (np.abs(rhs[interior]) <= tol).all())
    ok &= bool((np.abs(a[interior].sum(axis=1)) <= tol).all())
    return ok


# --------------------------------------------------------------------------
# a-7  solve: the assembled (non-symmetric) system, solved to tolerance.
# --------------------------------------------------------------------------

def assemble_csr(a: np.ndarray, shape):
    """CSR of the 7-point struct matrix a[N,7] (offsets C,-x,+x,-y,+y,-z,+z)."""
    import scipy.sparse as sp
    nz, ny, nx = shape
    N = nx * ny * nz
    strides = (0, -1, 1, -nx, nx, -nx * ny, nx * ny)
    rows, cols, vals = [], [], []
    m = np.arange(N)
    for s in range(7):
        v = a[:, s]
        sel = v != 0.0
        rows.append(m[sel])
        cols.append(m[sel] + strides[s])
        vals.append(v[sel])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))


def eliminate_dirichlet(A, rhs, x0, shape, mask, direction):
    """Move the identity rows (inactive + Dirichlet) out: SPD system on the
    active interior unknowns.  Algebraically identical to the reference's
    full system because x0 satisfies the identity rows exactly."""
    nz, ny, nx = shape
    n = shape[2 - direction]
    idx = np.arange(n).reshape([-1 if ax == 2 - direction else 1 for ax in range(3)])
    idx = np.broadcast_to(idx, shape).reshape(-1)
    m = mask.reshape(-1).astype(bool)
    diag = A.diagonal()
    unk = m & (idx != 0) & (idx != n - 1)
    # zero-diagonal-decoupled active cells are identity rows too
    offsum = np.abs(A).sum(axis=1).A1 - np.abs(diag) if hasattr(np.abs(A).sum(axis=1), "A1") else np.asarray(np.abs(A).sum(axis=1)).ravel() - np.abs(diag)
    unk &= offsum > 0
    fixed = ~unk
    xf = np.where(fixed, rhs, 0.0)      # identity rows: x = b
    Auu = A[unk][:, unk]
    bu = rhs[unk] - A[unk][:, fixed] @ xf[fixed]
    return Auu.tocsr(), bu, unk, xf


def reference_stop_norm(rhs: np.ndarray) -> float:
    """HYPRE FlexGMRES stops on ||r||_2 <= eps * ||b||_2 with b the full rhs
    (hypre v2.32.0 krylov/flexgmres.c; call site TortuosityHypre.cpp:667)."""
    return float(np.sqrt(np.dot(rhs, rhs)))


def solve_full(phase, phase_id, direction, vlo, vhi, eps=1e-12, dx=(1.0, 1.0, 1.0),
               maxiter=20000, mask=None):
    """Mask -> matrix -> eliminated SPD solve (Jacobi-PCG) -> full x field.
    Stops when ||r|| <= eps*||b_full|| (reference rule, SURVEY a-7)."""
    import scipy.sparse.linalg as spla
    if mask is None:
        mask = activity_mask(phase, phase_id, direction)
    a, rhs, x0 = fill_matrix(phase, mask, phase_id, direction, vlo, vhi, dx)
    A = assemble_csr(a, phase.shape)
    Auu, bu, unk, xf = eliminate_dirichlet(A, rhs, x0, phase.shape, mask, direction)
    bnorm = reference_stop_norm(rhs)
    x = xf.copy()
    iters = 0
    relres = 0.0
    if unk.any():
        d = Auu.diagonal()
        M = spla.LinearOperator(Auu.shape, matvec=lambda v: v / d)
        it = [0]

        def cb(_):
            it[0] += 1
        # scipy stops on its recurrence residual: aim at eps/2 so the TRUE residual
        # checked below is inside eps
        xu, info = spla.cg(Auu, bu, x0=x0[unk], rtol=0.0, atol=0.5 * eps * bnorm if bnorm > 0 else eps,
                           maxiter=maxiter, M=M, callback=cb)
        iters = it[0]
        x[unk] = xu
        relres = float(np.linalg.norm(bu - Auu @ xu) / (bnorm if bnorm > 0 else 1.0))
    return x.reshape(phase.shape), mask, dict(iters=iters, relres=relres, bnorm=bnorm)


def flexgmres(A, b, x0, precond, tol=1e-9, maxiter=200, k_dim=20, a_tol=0.0):
    """Restatement of HYPRE's FlexGMRES (hypre v2.32.0 krylov/flexgmres.c, the
    published right-preconditioned flexible GMRES(k) of Saad 1993) as the
    reference drives it (TortuosityHypre.cpp:666-688): tolerance ``tol``
    relative to ||b||_2 of the FULL rhs (||r0|| when b = 0), ``maxiter``
    Krylov steps in total, restart length k_dim = 20 (the reference never calls
    SetKDim, so HYPRE's default holds), modified Gram-Schmidt, Givens
    rotations, preconditioned directions z_j kept so x += sum y_j z_j, and the
    recurrence residual confirmed by a true residual b - A x before the solver
    reports convergence.  ``precond(v)`` applies M^-1 (HYPRE: one SMG V-cycle;
    any fixed or varying M is legal here).
    Returns (x, iterations, final_relative_residual, converged)."""
    x = np.array(x0, dtype=np.float64, copy=True)
    b_norm = float(np.sqrt(b @ b))
    r = b - A @ x
    r_norm = float(np.sqrt(r @ r))
    den = b_norm if b_norm > 0.0 else r_norm
    epsilon = max(a_tol, tol * den)
    it = 0
    converged = r_norm <= epsilon
    n = b.shape[0]
    while it < maxiter and not converged and r_norm > 0.0:
        P = np.empty((k_dim + 1, n))
        Z = np.empty((k_dim, n))
        hh = np.zeros((k_dim + 1, k_dim))
        c = np.zeros(k_dim)
        s = np.zeros(k_dim)
        rs = np.zeros(k_dim + 1)
        rs[0] = r_norm
        P[0] = r / r_norm
        i = 0
        while i < k_dim and it < maxiter:
            i += 1
            it += 1
            Z[i - 1] = precond(P[i - 1])
            w = A @ Z[i - 1]
            for j in range(i):                      # modified Gram-Schmidt
                hh[j, i - 1] = P[j] @ w
                w -= hh[j, i - 1] * P[j]
            t = float(np.sqrt(w @ w))
            hh[i, i - 1] = t
            if t != 0.0:
                w /= t
            P[i] = w
            for j in range(1, i):                   # earlier rotations on the new column
                t = hh[j - 1, i - 1]
                hh[j - 1, i - 1] = s[j - 1] * hh[j, i - 1] + c[j - 1] * t
                hh[j, i - 1] = -s[j - 1] * t + c[j - 1] * hh[j, i - 1]
            gamma = float(np.hypot(hh[i, i - 1], hh[i - 1, i - 1]))
            if gamma == 0.0:
                gamma = 1e-16
            c[i - 1] = hh[i - 1, i - 1] / gamma
            s[i - 1] = hh[i, i - 1] / gamma
            rs[i] = -s[i - 1] * rs[i - 1]
            rs[i - 1] = c[i - 1] * rs[i - 1]
            hh[i - 1, i - 1] = s[i - 1] * hh[i, i - 1] + c[i - 1] * hh[i - 1, i - 1]
            r_norm = abs(rs[i])
            if r_norm <= epsilon:
                break
        y = rs[:i].copy()                           # back substitution
        for k in range(i - 1, -1, -1):
            y[k] -= hh[k, k + 1:i] @ y[k + 1:i]
            y[k] /= hh[k, k]
        x += y @ Z[:i]
        r = b - A @ x                               # true residual at every restart
        r_norm = float(np.sqrt(r @ r))
        converged = r_norm <= epsilon
    relres = r_norm / den if den > 0.0 else 0.0
    return x, it, relres, bool(converged)


def solve_full_flexgmres(phase, phase_id, direction, vlo, vhi, eps=1e-9, dx=(1.0, 1.0, 1.0),
                         maxiter=200, mask=None, k_dim=20, drop_tol=1e-5, fill_factor=20.0):
    """The reference's OWN formulation of a-7: the assembled NON-symmetric
    system with its identity rows kept (no Dirichlet elimination), initial
    guess = the ramp of tortuosity_fillmtx, FlexGMRES(20) with the reference's
    tolerance / iteration cap (TortuosityHypre.cpp:142-143, 666-688) and
    m_converged = finite and relres <= eps (:687-688).  HYPRE's SMG V-cycle
    cannot be restated from this tree (un-vendored dependency); an incomplete
    LU of the same matrix stands in as the right preconditioner -- FlexGMRES
    accepts any M, and the converged x does not depend on it."""
    import scipy.sparse.linalg as spla
    if mask is None:
        mask = activity_mask(phase, phase_id, direction)
    a, rhs, x0 = fill_matrix(phase, mask, phase_id, direction, vlo, vhi, dx)
    A = assemble_csr(a, phase.shape)
    ilu = spla.spilu(A.tocsc(), drop_tol=drop_tol, fill_factor=fill_factor)
    x, iters, relres, conv = flexgmres(A, rhs, x0, ilu.solve, tol=eps, maxiter=maxiter, k_dim=k_dim)
    conv = bool(conv and np.isfinite(relres) and 0.0 <= relres <= eps)
    return x.reshape(phase.shape), mask, dict(iters=iters, relres=relres, converged=conv,
                                              bnorm=reference_stop_norm(rhs))


# --------------------------------------------------------------------------
# a-8 / a-9  global_fluxes + value  (TortuosityHypre.cpp:1000-1134, 761-891)
# --------------------------------------------------------------------------

def global_fluxes(x: np.ndarray, mask: np.ndarray, direction: int, dx=(1.0, 1.0, 1.0)):
    n = x.shape[2 - direction]
    m = mask.astype(bool)
    d = dx[direction]
    area = {0: dx[1] * dx[2], 1: dx[0] * dx[2], 2: dx[0] * dx[1]}[direction]
    if n < 2:
        return 0.0, 0.0, int(_plane(m, direction, 0).sum()), int(_plane(m, direction, n - 1).sum())
    b0, b1 = _plane(x, direction, 0), _plane(x, direction, 1)
    m0, m1 = _plane(m, direction, 0), _plane(m, direction, 1)
    fin = float(np.sum(np.where(m0 & m1, -(b1 - b0) / d, 0.0)))         # :1067-1083
    e0, e1 = _plane(x, direction, n - 1), _plane(x, direction, n - 2)
    me0, me1 = _plane(m, direction, n - 1), _plane(m, direction, n - 2)
    fout = float(np.sum(np.where(me0 & me1, -(e0 - e1) / d, 0.0)))      # :1086-1103
    return fin * area, fout * area, int(m0.sum()), int(me0.sum())


@dataclass
class TauResult:
    tau: float
    deff: float
    active_vf: float
    flux_in: float
    flux_out: float
    n_active: int
    converged: bool
    iters: int
    relres: float


def tau_from_fluxes(fin, fout, active_vf, shape, direction, vlo, vhi, dx=(1.0, 1.0, 1.0),
                    converged=True):
    """value() tail (TortuosityHypre.cpp:782-877) incl. NaN/Inf conventions."""
    eps = np.finfo(np.float64).eps
    tiny = 1e-15
    if active_vf <= eps or not converged:
        return float("nan"), 0.0
    mag_in, mag_out = abs(fin), abs(fout)
    avg = 0.5 * (mag_in + mag_out)
    if avg > tiny and abs(mag_in - mag_out) / avg > 1e-6:
        return float("nan"), 0.0
    nz, ny, nx = shape
    ext = (nx * dx[0], ny * dx[1], nz * dx[2])   # ProbLength = N*dx
    L = ext[direction]
    A = {0: ext[1] * ext[2], 1: ext[0] * ext[2], 2: ext[0] * ext[1]}[direction]
    grad = (vhi - vlo) / L
    if avg < tiny:
        return float("inf"), 0.0
    if abs(grad) < tiny:
        return float("inf"), 0.0
    deff = (avg / A) / abs(grad)
    if abs(deff) < tiny:
        return float("inf"), deff
    return active_vf / deff, deff


def tortuosity(phase, phase_id, direction, vlo=-1.0, vhi=1.0, eps=1e-12,
               dx=(1.0, 1.0, 1.0), method="pcg", maxiter=None) -> TauResult:
    """End-to-end oracle: mask -> solve -> flux -> tau.  method "pcg": the
    eliminated SPD system by Jacobi-PCG; "flexgmres": the reference's own
    non-symmetric system by FlexGMRES(20) (solve_full_flexgmres)."""
    mask = activity_mask(phase, phase_id, direction)
    n_active = int(mask.sum())
    active_vf = n_active / phase.size if phase.size else 0.0
    if active_vf <= np.finfo(np.float64).eps:
        return TauResult(float("nan"), 0.0, active_vf, 0.0, 0.0, n_active, False, 0, float("nan"))
    if method == "flexgmres":
        x, mask, info = solve_full_flexgmres(phase, phase_id, direction, vlo, vhi, eps, dx,
                                             maxiter=200 if maxiter is None else maxiter, mask=mask)
    else:
        x, mask, info = solve_full(phase, phase_id, direction, vlo, vhi, eps, dx, mask=mask,
                                   **({} if maxiter is None else {"maxiter": maxiter}))
    fin, fout, _, _ = global_fluxes(x, mask, direction, dx)
    conv = np.isfinite(info["relres"]) and info["relres"] <= eps
    tau, deff = tau_from_fluxes(fin, fout, active_vf, phase.shape, direction, vlo, vhi, dx, conv)
    return TauResult(tau, deff, active_vf, fin, fout, n_active, bool(conv), info["iters"], info["relres"])


# --------------------------------------------------------------------------
# synthetic workloads: random overlapping-sphere packing
# --------------------------------------------------------------------------

def sphere_packing(n: int, seed: int = 12345, radius: int = 12, solid_target: float = 0.60,
                   shape=None) -> np.ndarray:
    """uint8 {0,1}[z,y,x]: 1 = pore, 0 = solid.  Spheres (solid) with centres
    uniform in the box (PCG64(seed)) are added in batches until the solid
    fraction >= solid_target."""
    shp = (n, n, n) if shape is None else tuple(shape)
    rng = np.random.Generator(np.random.PCG64(seed))
    solid = np.zeros(shp, dtype=bool)
    r = int(radius)
    g = np.arange(-r, r + 1)
    ball = (g[:, None, None] ** 2 + g[None, :, None] ** 2 + g[None, None, :] ** 2) <= r * r
    vol = shp[0] * shp[1] * shp[2]
    batch = max(1, int(0.02 * vol / ball.sum()))
    while solid.mean() < solid_target:
        cs = np.stack([rng.integers(0, s, size=batch) for s in shp], axis=1)
        for cz, cy, cx in cs:
            z0, z1 = max(cz - r, 0), min(cz + r + 1, shp[0])
            y0, y1 = max(cy - r, 0), min(cy + r + 1, shp[1])
            x0, x1 = max(cx - r, 0), min(cx + r + 1, shp[2])
            solid[z0:z1, y0:y1, x0:x1] |= ball[z0 - cz + r:z1 - cz + r, y0 - cy + r:y1 - cy + r, x0 - cx + r:x1 - cx + r]
    return (~solid).astype(np.uint8)
