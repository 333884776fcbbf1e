"""Python mirror of the reference's homogenisation surface.

  OpenImpala::EffectiveDiffusivityHypre   src/props/EffectiveDiffusivityHypre.H
      ctor (geom, ba, dm, mf_phase_input, phase_id, dir_of_chi_k, solver_type,
            resultspath, verbose_level, write_plotfile)      .cpp:104-203
      bool solve()                                           .cpp:543-676
      getChiSolution(), getSolverConverged(), getSolverIterations(),
      getFinalRelativeResidualNorm()
  calculate_Deff_tensor_homogenization   src/props/Diffusion.cpp:60-167

All numerics run in the CUDA library through the C-ABI (problem =
OI_PROBLEM_CELL); the host only forms the 3x3 tensor from the per-direction sums
the device returns.
"""
from __future__ import annotations

import math

import numpy as np

from . import capi
from .tortuosity import Direction, ParmParse, SolverType


class EffectiveDiffusivityHypre:
    """One corrector problem chi_k on a periodic box; phase == phase_id cells are
    the pore space (D = 1), everything else is solid (D = 0)."""

    SolverType = SolverType

    def __init__(self, geom, ba, dm, mf_phase_input: np.ndarray, phase_id: int, dir_of_chi_k: Direction,
                 solver_type: SolverType, resultspath: str = "", verbose_level: int = 0,
                 write_plotfile: bool = False, **b200):
        geom = geom or {}
        self._dx = tuple(float(v) for v in geom.get("dx", (1.0, 1.0, 1.0)))
        self._phase_field = np.ascontiguousarray(mf_phase_input)
        if self._phase_field.ndim != 3:
            raise ValueError("phase field must be 3-D [z, y, x]")
        self._phase_id, self._dir = int(phase_id), Direction(dir_of_chi_k)
        self._solvertype = SolverType(solver_type)
        if self._solvertype != SolverType.FlexGMRES:                     # .cpp:616-619
            raise ValueError("Unsupported solver type requested in EffectiveDiffusivityHypre::solve: "
                             f"{int(self._solvertype)}")
        self._eps = ParmParse.query("hypre.eps", 1e-9)                   # .cpp:124, 156-158
        self._maxiter = ParmParse.query("hypre.maxiter", 1000)
        if not self._eps > 0.0:
            raise ValueError("Solver tolerance (eps) must be positive")
        if not self._maxiter > 0:
            raise ValueError("Solver max iterations must be positive")
        for d in self._dx:
            if not d > 0.0:
                raise ValueError("Cell size must be positive.")
        self._verbose = int(verbose_level)
        self._num_iterations = -1
        self._final_res_norm = math.nan
        self._converged = False
        # multi-GPU: mf_phase_input is this rank's z-slab, `global_shape` the whole (periodic) box
        shape = b200.pop("global_shape", None) or self._phase_field.shape
        self._solver = capi.Solver(shape, int(self._dir), self._phase_id, 0.0, 1.0,
                                   eps=self._eps, maxiter=self._maxiter, dx=self._dx, verbose=self._verbose,
                                   problem=capi.OI_PROBLEM_CELL, **b200)
        self._solver.set_phase(self._phase_field)
        self._n_active = self._solver.build_mask()
        if self._n_active == 0:                                          # .cpp:186-196
            self._converged = True
            self._num_iterations = 0
            self._final_res_norm = 0.0

    def solve(self) -> bool:
        if self._n_active == 0:                                          # .cpp:559-572
            self._converged, self._num_iterations, self._final_res_norm = True, 0, 0.0
            return True
        info = self._solver.solve()
        self.last_info = info
        self._num_iterations = info.iterations
        self._final_res_norm = info.rel_residual
        ok = not (math.isnan(info.rel_residual) or math.isinf(info.rel_residual))
        self._converged = bool(ok and info.converged)                    # .cpp:607-608
        return self._converged

    def getChiSolution(self) -> np.ndarray:
        """chi_k on the box / this rank's slab (zero in the solid; zero everywhere if not converged,
        .cpp:631-637)."""
        if not self._converged or self._n_active == 0:
            return np.zeros(self._phase_field.shape)
        return self._solver.solution()

    def gradient_sums(self):
        """sum over pore cells of d chi_k/dx_a (a = x, y, z), evaluated on the device."""
        if not self._converged or self._n_active == 0:
            return (0.0, 0.0, 0.0)
        return self._solver.cell_gradient_sums()[0]

    def getSolverConverged(self): return self._converged
    def getSolverIterations(self): return self._num_iterations
    def getFinalRelativeResidualNorm(self): return self._final_res_norm
    def getNumActiveCells(self): return self._n_active

    @property
    def solver(self):
        return self._solver

    def close(self):
        self._solver.close()


def calculate_Deff_tensor_homogenization(mf_phase: np.ndarray, phase_id: int, solver_type=SolverType.FlexGMRES,
                                         geom=None, verbose: int = 0, **b200):
    """The three corrector solves of Diffusion.cpp:511-589 and the tensor of
    Diffusion.cpp:60-167.  Returns (D[3][3] as ndarray, all_converged, per-direction info)."""
    gshape = b200.get("global_shape", None) or mf_phase.shape
    n_total = int(np.prod(gshape))
    D = np.zeros((3, 3))
    infos = []
    all_ok = True
    for k in (Direction.X, Direction.Y, Direction.Z):
        s = EffectiveDiffusivityHypre(geom, None, None, mf_phase, phase_id, k, solver_type, "", verbose, False, **b200)
        try:
            ok = s.solve()
            infos.append(dict(direction=k.name, converged=ok, iterations=s.getSolverIterations(),
                              rel_residual=s.getFinalRelativeResidualNorm()))
            if not ok:
                all_ok = False
                break                                                    # Diffusion.cpp:546-549
            sums = s.gradient_sums()
            n_act = s.getNumActiveCells()
            for a in range(3):
                D[a][int(k)] = ((n_act if a == int(k) else 0.0) - sums[a]) / n_total
        finally:
            s.close()
    if not all_ok:
        D[:] = math.nan
    return D, all_ok, infos
