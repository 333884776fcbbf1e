"""Python mirror of the reference's class surface for the tortuosity path.

Same names, argument meaning and error behaviour as
  OpenImpala::Direction        src/props/Tortuosity.H:9-13
  OpenImpala::VolumeFraction   src/props/VolumeFraction.H:21-91
  OpenImpala::TortuosityHypre  src/props/TortuosityHypre.H:36-186
so that the parity tests read like src/props/tTortuosity.cpp.  All numerics run
in the CUDA library through the C-ABI (openimpala_b200.capi); the only host
arithmetic is the scalar tail of value() (TortuosityHypre.cpp:782-877).

Phase fields are numpy arrays indexed [z, y, x] (x fastest), i.e. the dense
equivalent of the reference's single-component iMultiFab.
"""
from __future__ import annotations

import enum
import math
import sys

import numpy as np

from . import capi


class Direction(enum.IntEnum):
    X = 0
    Y = 1
    Z = 2


class SolverType(enum.IntEnum):
    """TortuosityHypre::SolverType, src/props/TortuosityHypre.H:42-50 (order
    matters: Diffusion.cpp:664-665 static_casts across enums).  The reference
    snapshot aborts on everything but FlexGMRES (TortuosityHypre.cpp:695-697);
    here every value selects the same multigrid-preconditioned CG because the
    converged answer is solver independent."""
    Jacobi = 0
    GMRES = 1
    FlexGMRES = 2
    PCG = 3
    BiCGSTAB = 4
    SMG = 5
    PFMG = 6


def string_to_solver_type(s: str) -> SolverType:
    """Diffusion.cpp:45-58 (case-insensitive; unknown -> abort)."""
    table = {t.name.lower(): t for t in SolverType}
    try:
        return table[s.lower()]
    except KeyError:
        raise ValueError(f"Invalid solver string: '{s}'.")


class ParmParse:
    """The few amrex::ParmParse queries the class itself performs
    (TortuosityHypre.cpp:147-151, 255-256): a process-global table."""
    table: dict = {}

    @classmethod
    def query(cls, key, default):
        return type(default)(cls.table[key]) if key in cls.table else default


class VolumeFraction:
    """VolumeFraction(fm, phase=0, comp=0), src/props/VolumeFraction.H:37."""

    def __init__(self, fm: np.ndarray, phase: int = 0, comp: int = 0):
        if comp != 0:
            raise ValueError("VolumeFraction: Component index out of bounds.")  # .cpp:17-18
        self._fm = fm           # holds a reference, like the iMultiFab&
        self._phase = int(phase)

    def value(self, local: bool = False):
        """-> (phase_count, total_count), VolumeFraction.cpp:22-66."""
        return capi.count_phase(self._fm, self._phase)

    def value_vf(self, local: bool = False) -> float:
        pc, tc = self.value(local)
        return pc / tc if tc > 0 else 0.0       # VolumeFraction.H:68-72


_TINY = 1.0e-15                                   # tiny_flux_threshold, TortuosityHypre.cpp:63
_EPS = sys.float_info.epsilon


def tau_from_fluxes(flux_in, flux_out, active_vf, length, area, vlo, vhi):
    """Scalar tail of TortuosityHypre::value() (TortuosityHypre.cpp:794-877).
    Returns (tau, deff, flux_conserved)."""
    mag_in, mag_out = abs(flux_in), abs(flux_out)
    avg = 0.5 * (mag_in + mag_out)
    conserved = True
    if avg > _TINY and abs(mag_in - mag_out) / avg > 1.0e-6:     # :800-804
        conserved = False
    if not conserved:
        return math.nan, 0.0, False                              # :819-823
    grad = (vhi - vlo) / length                                  # :841
    if avg < _TINY:                                              # :846-851
        return (math.inf if active_vf > _EPS else math.nan), 0.0, True
    if active_vf <= _EPS:                                        # :854-858
        return math.nan, 0.0, True
    if abs(grad) < _TINY:                                        # :860-864
        return math.inf, 0.0, True
    deff = (avg / area) / abs(grad)                              # :868
    if abs(deff) < _TINY:                                        # :869-873
        return math.inf, deff, True
    return active_vf / deff, deff, True                          # :876


class TortuosityHypre:
    """Constructor arguments follow src/props/TortuosityHypre.H:68-80 with the
    AMReX containers replaced by their dense content: `geom` is a dict with
    optional 'dx' (cell size triple; ProbLength = N*dx), ba/dm are accepted and
    ignored (single box per GPU slab)."""

    SolverType = SolverType

    def __init__(self, geom, ba, dm, mf_phase_input: np.ndarray, vf: float, phase: int,
                 dir: Direction, solvertype: SolverType, resultspath: str, vlo: float = 0.0,
                 vhi: float = 1.0, verbose: int = 0, write_plotfile: bool = False, **b200):
        geom = geom or {}
        self._dx = tuple(float(v) for v in geom.get("dx", (1.0, 1.0, 1.0)))
        self._phase_field = np.ascontiguousarray(mf_phase_input)   # deep copy, .cpp:132
        if self._phase_field.ndim != 3:
            raise ValueError("phase field must be 3-D [z, y, x]")
        self._vf, self._phase, self._dir = float(vf), int(phase), Direction(dir)
        self._solvertype = SolverType(solvertype)
        self._resultspath, self._vlo, self._vhi = resultspath, float(vlo), float(vhi)
        self._verbose, self._write_plotfile = int(verbose), bool(write_plotfile)
        # defaults + ParmParse overrides, TortuosityHypre.cpp:142-151
        self._eps = ParmParse.query("hypre.eps", 1e-9)
        self._maxiter = ParmParse.query("hypre.maxiter", 200)
        self._verbose = ParmParse.query("tortuosity.verbose", self._verbose)
        if not (0.0 <= self._vf <= 1.0):
            raise ValueError("Original Volume fraction must be between 0 and 1")   # :158
        if not self._eps > 0.0:
            raise ValueError("Solver tolerance (eps) must be positive")             # :159
        if not self._maxiter > 0:
            raise ValueError("Solver max iterations must be positive")              # :160
        self._value = math.nan
        self._first_call = True
        self._num_iterations = -1
        self._final_res_norm = math.nan
        self._converged = False
        self._flux_in = self._flux_out = 0.0
        self._active_vf = 0.0
        self.last_info = None

        # multi-GPU: mf_phase_input is this rank's z-slab, `global_shape` the whole box
        shape = b200.pop("global_shape", None) or self._phase_field.shape
        self._solver = capi.Solver(shape, int(self._dir), self._phase, self._vlo,
                                   self._vhi, eps=self._eps, maxiter=self._maxiter, dx=self._dx,
                                   verbose=self._verbose, **b200)
        self._solver.set_phase(self._phase_field)
        self._solver.remspot(ParmParse.query("tortuosity.remspot_passes", 0))     # :248-292
        n_active = self._solver.build_mask()                                       # :394-558
        total = int(np.prod(self._solver.global_shape))   # Geometry::Domain().numPts()
        self._n_active = n_active
        self._active_vf = (n_active / total) if total > 0 else 0.0                  # :552-553
        if self._active_vf <= _EPS:                                                 # :170-178
            self._first_call = False
            self._value = math.nan

    # ---- getters, TortuosityHypre.H:114-121
    def getSolverConverged(self): return self._converged
    def getFinalRelativeResidualNorm(self): return self._final_res_norm
    def getSolverIterations(self): return self._num_iterations
    def getFluxIn(self): return self._flux_in
    def getFluxOut(self): return self._flux_out
    def getActiveVolumeFraction(self): return self._active_vf

    def checkMatrixProperties(self) -> bool:                                        # :896-982
        if self._active_vf <= _EPS:
            return True
        return self._solver.check_matrix_properties()

    def _solve(self) -> bool:                                                       # :654-756
        info = self._solver.solve()
        self.last_info = info
        self._num_iterations = info.iterations
        self._final_res_norm = info.rel_residual
        ok = not (math.isnan(info.rel_residual) or math.isinf(info.rel_residual))
        self._converged = bool(ok and info.converged)
        return self._converged

    def value(self, refresh: bool = False) -> float:                                # :761-891
        if self._active_vf <= _EPS and not self._first_call:
            return math.nan
        if self._first_call or refresh:
            if self._active_vf <= _EPS:
                self._value = math.nan
                self._first_call = False
                return self._value
            if not self._solve():
                self._value = math.nan
                self._first_call = False
                return self._value
            self._flux_in, self._flux_out, _, _ = self._solver.fluxes()             # :790
            nz, ny, nx = self._solver.global_shape
            ext = (nx * self._dx[0], ny * self._dx[1], nz * self._dx[2])            # ProbLength
            d = int(self._dir)
            length = ext[d]
            area = ext[1] * ext[2] if d == 0 else (ext[0] * ext[2] if d == 1 else ext[0] * ext[1])
            self._value, self._deff, _ = tau_from_fluxes(self._flux_in, self._flux_out,
                                                         self._active_vf, length, area,
                                                         self._vlo, self._vhi)
        self._first_call = False
        return self._value

    # conveniences beyond the reference surface
    def solution(self):
        return self._solver.solution()

    def active_mask(self):
        return self._solver.mask()

    @property
    def solver(self):
        return self._solver

    def close(self):
        self._solver.close()
