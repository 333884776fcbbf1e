"""ctypes binding of include/openimpala_b200.h (the C-ABI shared library).

This is the binding a Python caller of the reference would add; the C++ host
classes in openimpala_b200/host call the same entry points directly.  There is
no CPU fallback: if the library is missing, or no CUDA device is visible, the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libopenimpala_b200.so")

OI_OK = 0
OI_PRECOND_MG, OI_PRECOND_JACOBI = 0, 1
OI_HALO_AUTO, OI_HALO_NCCL, OI_HALO_PEER = 0, 1, 2
OI_PROBLEM_TORTUOSITY, OI_PROBLEM_CELL = 0, 1


class OiError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"openimpala_b200 error {code}: {msg}")
        self.code = code


class oi_params(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("z_begin", C.c_int32), ("nz_local", C.c_int32),
        ("direction", C.c_int32), ("phase_id", C.c_int32),
        ("vlo", C.c_double), ("vhi", C.c_double),
        ("dx", C.c_double * 3),
        ("eps", C.c_double),
        ("maxiter", C.c_int32), ("verbose", C.c_int32), ("device", C.c_int32),
        ("precond", C.c_int32), ("mg_degree", C.c_int32), ("stencil_variant", C.c_int32),
        ("flux_polish", C.c_int32), ("halo_mode", C.c_int32), ("problem", C.c_int32),
        ("comm", C.c_void_p),
    ]


class oi_solve_info(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("converged", C.c_int32),
        ("rel_residual", C.c_double), ("b_norm", C.c_double),
        ("solve_ms", C.c_double), ("setup_ms", C.c_double),
    ]


# every symbol include/openimpala_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SIGS = {
    "oi_version": (C.c_int, []),
    "oi_last_error": (C.c_char_p, []),
    "oi_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "oi_default_params": (None, [C.POINTER(oi_params)]),
    "oi_comm_unique_id": (C.c_int, [_P]),
    "oi_comm_create": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, _P, C.c_int32]),
    "oi_comm_destroy": (C.c_int, [_P]),
    "oi_comm_allreduce_sum_i64": (C.c_int, [_P, C.POINTER(C.c_int64), C.c_int32]),
    "oi_count_phase_i32": (C.c_int, [_P, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "oi_count_phase_u8": (C.c_int, [_P, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "oi_create": (C.c_int, [C.POINTER(_P), C.POINTER(oi_params)]),
    "oi_destroy": (C.c_int, [_P]),
    "oi_set_phase_i32": (C.c_int, [_P, _P]),
    "oi_set_phase_u8": (C.c_int, [_P, _P]),
    "oi_set_phase_device_u8": (C.c_int, [_P, _P]),
    "oi_phase_stream_begin": (C.c_int, [_P, C.c_int32]),
    "oi_phase_stream_buffer": (C.c_int, [_P, C.c_int32, C.POINTER(C.POINTER(C.c_uint8))]),
    "oi_phase_stream_submit": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "oi_phase_stream_end": (C.c_int, [_P]),
    "oi_volume_fraction": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "oi_remspot": (C.c_int, [_P, C.c_int32]),
    "oi_build_mask": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "oi_solve": (C.c_int, [_P, C.POINTER(oi_solve_info)]),
    "oi_fluxes": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double),
                            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "oi_cell_gradient_sums": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "oi_check_matrix_properties": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "oi_get_mask_u8": (C.c_int, [_P, _P]),
    "oi_get_solution": (C.c_int, [_P, _P]),
    "oi_set_solution": (C.c_int, [_P, _P]),
    "oi_get_initial_guess": (C.c_int, [_P, _P]),
    "oi_get_rhs": (C.c_int, [_P, _P]),
    "oi_get_matrix_rows": (C.c_int, [_P, _P]),
    "oi_apply_operator": (C.c_int, [_P, _P, _P]),
    "oi_apply_precond": (C.c_int, [_P, _P, _P]),
    "oi_time_kernel": (C.c_int, [_P, C.c_char_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "oi_timer_record": (C.c_int, [_P, C.c_int32]),
    "oi_timer_elapsed_ms": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "oi_release_cached_memory": (C.c_int, [C.POINTER(C.c_int64)]),
    "oi_sparsity": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "oi_halo_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "oi_graph_info": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "oi_launch_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    # the reference's own bind(c) entry points (Fortran convention: everything by reference)
    "tortuosity_fillmtx": (None, [_P] * 20),
    "tortuosity_remspot": (None, [_P] * 8),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OiError(-1, f"{LIB_PATH} not built; run `python -m openimpala_b200.build`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(code: int):
    if code != OI_OK:
        raise OiError(code, load().oi_last_error().decode())


def device_count() -> int:
    n = C.c_int(0)
    _check(load().oi_device_count(C.byref(n)))
    return n.value


def default_params() -> oi_params:
    p = oi_params()
    load().oi_default_params(C.byref(p))
    return p


def count_phase(field: np.ndarray, phase: int):
    """VolumeFraction::value on a host field -> (phase_count, total_count)."""
    lib = load()
    pc, tc = C.c_int64(0), C.c_int64(0)
    a = np.ascontiguousarray(field)
    if a.dtype == np.uint8:
        _check(lib.oi_count_phase_u8(a.ctypes.data, a.size, phase, C.byref(pc), C.byref(tc)))
    else:
        a = np.ascontiguousarray(a, dtype=np.int32)
        _check(lib.oi_count_phase_i32(a.ctypes.data, a.size, phase, C.byref(pc), C.byref(tc)))
    return pc.value, tc.value


def release_cached_memory() -> int:
    """Return idle cached device blocks to the driver; -> bytes released."""
    b = C.c_int64(0)
    _check(load().oi_release_cached_memory(C.byref(b)))
    return b.value


def comm_unique_id() -> bytes:
    """128-byte ncclUniqueId (call on one rank, hand the bytes to the others)."""
    buf = C.create_string_buffer(128)
    _check(load().oi_comm_unique_id(C.cast(buf, _P)))
    return buf.raw


class Comm:
    """z-slab communicator: rank r owns slab r (stands in for MPI_COMM_WORLD)."""

    def __init__(self, rank: int, n_ranks: int, unique_id: bytes, device: int = -1):
        self._lib = load()
        self._h = _P(None)
        self.rank, self.n_ranks = int(rank), int(n_ranks)
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(self._lib.oi_comm_create(C.byref(self._h), self.rank, self.n_ranks, C.cast(buf, _P), device))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            self._lib.oi_comm_destroy(self._h)
            self._h = _P(None)


def slab_partition(nz: int, n_ranks: int, align: int = 0):
    """Contiguous z-slabs [(z_begin, nz_local)] whose boundaries are multiples of
    `align` planes (so 2x2x2 multigrid aggregates never straddle ranks).
    align = 0 picks the largest power of two <= nz // n_ranks, capped at 64."""
    if n_ranks <= 1:
        return [(0, nz)]
    if align <= 0:
        align = 1
        while align * 2 <= max(1, nz // n_ranks) and align < 64:
            align *= 2
    blocks = nz // align
    if blocks < n_ranks:
        raise ValueError(f"nz={nz} too small for {n_ranks} slabs aligned to {align}")
    out, z = [], 0
    for r in range(n_ranks):
        nb = blocks // n_ranks + (1 if r < blocks % n_ranks else 0)
        n = nb * align if r < n_ranks - 1 else nz - z
        out.append((z, n))
        z += n
    return out


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def ref_tortuosity_fillmtx(p, p_lo, active_mask, mask_lo, bxlo, bxhi, domlo, domhi, dxinv, vlo, vhi, phase, direction,
                           xinit=None):
    """tortuosity_fillmtx exactly as the reference's C++ calls it (TortuosityHypre.cpp:612-632): `p` and
    `active_mask` are int32 arrays [z, y, x] whose element [0, 0, 0] has index p_lo / mask_lo (ghost cells
    included); returns (a[n, 7], rhs[n], xinit[n]) for the cells of bxlo..bxhi, x fastest."""
    lib = load()
    p = np.ascontiguousarray(p, dtype=np.int32)
    m = np.ascontiguousarray(active_mask, dtype=np.int32)
    p_hi = [p_lo[d] + p.shape[2 - d] - 1 for d in range(3)]
    m_hi = [mask_lo[d] + m.shape[2 - d] - 1 for d in range(3)]
    n = int(np.prod([bxhi[d] - bxlo[d] + 1 for d in range(3)]))
    a = np.full((n, 7), np.nan)
    rhs = np.full(n, np.nan)
    x = np.zeros(n) if xinit is None else np.ascontiguousarray(xinit, dtype=np.float64).copy()
    dxi = (C.c_double * 3)(*[float(v) for v in dxinv])
    lib.tortuosity_fillmtx(a.ctypes.data, rhs.ctypes.data, x.ctypes.data, C.byref(C.c_int(n)), p.ctypes.data,
                           _i3(p_lo), _i3(p_hi), m.ctypes.data, _i3(mask_lo), _i3(m_hi), _i3(bxlo), _i3(bxhi),
                           _i3(domlo), _i3(domhi), dxi, C.byref(C.c_double(vlo)), C.byref(C.c_double(vhi)),
                           C.byref(C.c_int(phase)), C.byref(C.c_int(direction)), C.byref(C.c_int(0)))
    return a, rhs, x


def ref_tortuosity_remspot(q, q_lo, bxlo, bxhi, domlo, domhi):
    """tortuosity_remspot as the reference calls it (TortuosityHypre.cpp:270-290): one in-place pass over
    bxlo..bxhi of the int32 array q [z, y, x] whose element [0, 0, 0] has index q_lo.  Returns the filtered copy."""
    lib = load()
    q = np.ascontiguousarray(q, dtype=np.int32).copy()
    q_hi = [q_lo[d] + q.shape[2 - d] - 1 for d in range(3)]
    lib.tortuosity_remspot(q.ctypes.data, _i3(q_lo), _i3(q_hi), C.byref(C.c_int(1)), _i3(bxlo), _i3(bxhi),
                           _i3(domlo), _i3(domhi))
    return q


class Solver:
    """Thin RAII wrapper of one oi_solver handle (one image / phase / direction)."""

    def __init__(self, shape, direction: int, phase_id: int = 1, vlo: float = 0.0, vhi: float = 1.0,
                 eps: float = 1e-9, maxiter: int = 200, dx=(1.0, 1.0, 1.0), precond: int = OI_PRECOND_MG,
                 mg_degree: int = 0, stencil_variant: int = 0, flux_polish: int = 0, device: int = -1,
                 z_begin: int = 0, nz_local: int = 0, comm: "Comm | None" = None, verbose: int = 0,
                 halo_mode: int = OI_HALO_AUTO, problem: int = OI_PROBLEM_TORTUOSITY):
        self._lib = load()
        self._h = _P(None)
        nz, ny, nx = (int(s) for s in shape)
        p = default_params()
        p.nx, p.ny, p.nz = nx, ny, nz
        p.z_begin, p.nz_local = z_begin, (nz_local if nz_local > 0 else nz)
        p.direction, p.phase_id = int(direction), int(phase_id)
        p.vlo, p.vhi = float(vlo), float(vhi)
        p.dx[0], p.dx[1], p.dx[2] = (float(d) for d in dx)
        p.eps, p.maxiter, p.verbose, p.device = float(eps), int(maxiter), int(verbose), int(device)
        p.precond, p.mg_degree, p.stencil_variant = int(precond), int(mg_degree), int(stencil_variant)
        p.flux_polish = int(flux_polish)
        p.halo_mode = int(halo_mode)
        p.problem = int(problem)
        self._comm = comm                       # keep the communicator alive
        p.comm = comm.handle if comm is not None else None
        self.params = p
        self.global_shape = (nz, ny, nx)
        self.local_shape = (p.nz_local, ny, nx)
        _check(self._lib.oi_create(C.byref(self._h), C.byref(p)))

    # -- lifetime
    def close(self):
        if self._h:
            self._lib.oi_destroy(self._h)
            self._h = _P(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- setup
    def set_phase(self, phase: np.ndarray):
        a = np.ascontiguousarray(phase)
        if a.shape != self.local_shape:
            raise ValueError(f"phase shape {a.shape} != local slab {self.local_shape}")
        if a.dtype == np.uint8:
            _check(self._lib.oi_set_phase_u8(self._h, a.ctypes.data))
        else:
            a = np.ascontiguousarray(a, dtype=np.int32)
            _check(self._lib.oi_set_phase_i32(self._h, a.ctypes.data))

    def set_phase_streamed(self, read_planes, planes_per_chunk: int = 16):
        """Streamed upload: read_planes(z0, nz, out) fills out[nz, ny, nx] (uint8, a view of a
        pinned staging buffer) with planes z0 .. z0+nz-1 of the local slab; decoding of chunk
        k+1 overlaps the upload of chunk k."""
        nz, ny, nx = self.local_shape
        chunk = max(1, min(int(planes_per_chunk), nz))
        _check(self._lib.oi_phase_stream_begin(self._h, chunk))
        which = 0
        for z0 in range(0, nz, chunk):
            n = min(chunk, nz - z0)
            buf = C.POINTER(C.c_uint8)()
            _check(self._lib.oi_phase_stream_buffer(self._h, which, C.byref(buf)))
            view = np.ctypeslib.as_array(buf, shape=(chunk, ny, nx))
            read_planes(z0, n, view[:n])
            _check(self._lib.oi_phase_stream_submit(self._h, which, z0, n))
            which ^= 1
        _check(self._lib.oi_phase_stream_end(self._h))

    def set_phase_device(self, dev_ptr: int):
        _check(self._lib.oi_set_phase_device_u8(self._h, _P(dev_ptr)))

    def volume_fraction(self):
        pc, tc = C.c_int64(0), C.c_int64(0)
        _check(self._lib.oi_volume_fraction(self._h, C.byref(pc), C.byref(tc)))
        return pc.value, tc.value

    def remspot(self, passes: int):
        _check(self._lib.oi_remspot(self._h, int(passes)))

    def build_mask(self) -> int:
        n = C.c_int64(0)
        _check(self._lib.oi_build_mask(self._h, C.byref(n)))
        return n.value

    # -- solve
    def solve(self) -> oi_solve_info:
        info = oi_solve_info()
        _check(self._lib.oi_solve(self._h, C.byref(info)))
        return info

    def fluxes(self):
        fi, fo = C.c_double(0), C.c_double(0)
        ni, no = C.c_int64(0), C.c_int64(0)
        _check(self._lib.oi_fluxes(self._h, C.byref(fi), C.byref(fo), C.byref(ni), C.byref(no)))
        return fi.value, fo.value, ni.value, no.value

    def cell_gradient_sums(self):
        """Cell problem: (sum_active d chi/dx, d chi/dy, d chi/dz), n_active."""
        sums = (C.c_double * 3)()
        n = C.c_int64(0)
        _check(self._lib.oi_cell_gradient_sums(self._h, sums, C.byref(n)))
        return (sums[0], sums[1], sums[2]), n.value

    def check_matrix_properties(self) -> bool:
        ok = C.c_int32(0)
        _check(self._lib.oi_check_matrix_properties(self._h, C.byref(ok)))
        return bool(ok.value)

    # -- read-backs
    def _out(self, fn, dtype, extra=()):
        a = np.empty(self.local_shape + tuple(extra), dtype=dtype)
        _check(fn(self._h, a.ctypes.data))
        return a

    def mask(self):
        return self._out(self._lib.oi_get_mask_u8, np.uint8)

    def solution(self):
        return self._out(self._lib.oi_get_solution, np.float64)

    def set_solution(self, x: np.ndarray):
        a = np.ascontiguousarray(x, dtype=np.float64)
        _check(self._lib.oi_set_solution(self._h, a.ctypes.data))

    def initial_guess(self):
        return self._out(self._lib.oi_get_initial_guess, np.float64)

    def rhs(self):
        return self._out(self._lib.oi_get_rhs, np.float64)

    def matrix_rows(self):
        return self._out(self._lib.oi_get_matrix_rows, np.float64, (7,))

    def apply_operator(self, x: np.ndarray):
        a = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.local_shape, dtype=np.float64)
        _check(self._lib.oi_apply_operator(self._h, a.ctypes.data, y.ctypes.data))
        return y

    def apply_precond(self, r: np.ndarray):
        a = np.ascontiguousarray(r, dtype=np.float64)
        z = np.empty(self.local_shape, dtype=np.float64)
        _check(self._lib.oi_apply_precond(self._h, a.ctypes.data, z.ctypes.data))
        return z

    def time_kernel(self, name: str, reps: int = 20):
        ms, cells = C.c_double(0), C.c_int64(0)
        _check(self._lib.oi_time_kernel(self._h, name.encode(), reps, C.byref(ms), C.byref(cells)))
        return ms.value, cells.value

    def timer_record(self, slot: int):
        _check(self._lib.oi_timer_record(self._h, slot))

    def timer_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double(0)
        _check(self._lib.oi_timer_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def sparsity(self):
        """(unknowns, 2-cell groups with an unknown, 4-cell groups with an unknown) of the local slab."""
        a = (C.c_int64 * 3)()
        _check(self._lib.oi_sparsity(self._h, a))
        return a[0], a[1], a[2]

    def halo_info(self):
        """(oi_halo_mode in use, ghost-plane exchanges done through peer memory)."""
        m, n = C.c_int32(0), C.c_int64(0)
        _check(self._lib.oi_halo_info(self._h, C.byref(m), C.byref(n)))
        return m.value, n.value

    def graph_info(self):
        """(iterations replayed as a CUDA graph so far, kernel nodes per captured iteration)."""
        r, k = C.c_int64(0), C.c_int64(0)
        _check(self._lib.oi_graph_info(self._h, C.byref(r), C.byref(k)))
        return r.value, k.value

    def launch_count(self) -> int:
        n = C.c_int64(0)
        _check(self._lib.oi_launch_count(self._h, C.byref(n)))
        return n.value
