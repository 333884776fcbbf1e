"""GPU parity tests of the homogenisation cell problem:
the CUDA path (problem = OI_PROBLEM_CELL) through the C-ABI against the numpy
restatement of effdiff_fillmtx / EffectiveDiffusivityHypre /
calculate_Deff_tensor_homogenization (oracle/oi_effdiff.py).

Rows and right-hand sides bit-exact; operator apply to 1e-12; corrector field and
D_eff tensor within 1e-6 of the oracle's 1e-12 solve (north_star tolerance).
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

DEFF_TOL = 1e-6


def _blobs(shape, seed, porosity=0.5, sigma=1.5):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    f = ndimage.gaussian_filter(rng.standard_normal(shape), sigma, mode="wrap")
    return (f > np.quantile(f, 1.0 - porosity)).astype(np.int32)


@pytest.fixture(scope="module")
def capi(built_lib):
    from openimpala_b200 import capi as c
    assert c.device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return c


# shapes: ring kernels (nx % 4 == 0, full and partial 64-wide tiles), z-march fallback (odd nx),
# tiny periodic extents (2 and 3 cells), non-cubic boxes
CASES = [((12, 16, 20), 1, 0.55), ((9, 11, 13), 2, 0.6), ((16, 24, 64), 3, 0.5), ((10, 20, 100), 4, 0.45),
         ((3, 5, 8), 5, 0.7), ((2, 4, 4), 6, 0.7), ((33, 18, 68), 7, 0.5), ((20, 17, 36), 8, 0.4)]


@pytest.mark.parametrize("shape,seed,por", CASES)
@pytest.mark.parametrize("direction", [0, 1, 2])
def test_rows_rhs_and_operator(capi, shape, seed, por, direction):
    from oracle import oi_effdiff as oe
    ph = _blobs(shape, seed, por)
    dx = (1.0, 1.0, 1.0) if seed % 2 else (0.5, 1.25, 2.0)
    for phase_id in (1, 0):
        a, rhs, _ = oe.fill_matrix(ph, phase_id, direction, dx)
        for variant in (0, 1):
            with capi.Solver(shape, direction, phase_id, dx=dx, problem=capi.OI_PROBLEM_CELL,
                             stencil_variant=variant) as s:
                s.set_phase(ph)
                n_active = s.build_mask()
                assert n_active == int((ph == phase_id).sum())
                assert np.array_equal(s.mask().astype(bool), ph == phase_id)
                assert np.array_equal(s.matrix_rows().reshape(-1, 7), a)     # coefficients are exact
                assert np.array_equal(s.rhs().ravel(), rhs)                   # same operation order
                assert s.check_matrix_properties()
                A = oe.assemble_csr_periodic(a, shape)
                act = (ph == phase_id).ravel()
                rng = np.random.default_rng(seed)
                x = np.where(act, rng.standard_normal(act.size), 0.0)
                y_ref = np.where(act, A @ x, 0.0)
                y = s.apply_operator(x.reshape(shape)).ravel()
                np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("shape,seed,por", [((24, 24, 24), 11, 0.5), ((16, 20, 36), 12, 0.6), ((21, 15, 18), 13, 0.45)])
def test_chi_and_tensor_match_oracle(capi, shape, seed, por):
    from openimpala_b200.effdiff import EffectiveDiffusivityHypre, calculate_Deff_tensor_homogenization
    from openimpala_b200.tortuosity import Direction, SolverType
    from oracle import oi_effdiff as oe
    ph = _blobs(shape, seed, por)
    act = ph == 1
    for k in (Direction.X, Direction.Y, Direction.Z):
        chi_ref, _, rel = oe.solve_chi(ph, 1, int(k), eps=1e-12)
        assert rel < 1e-10
        s = EffectiveDiffusivityHypre(None, None, None, ph, 1, k, SolverType.FlexGMRES)
        assert s.solve() and s.getFinalRelativeResidualNorm() <= 1e-9
        assert 0 < s.getSolverIterations() < 60
        chi = s.getChiSolution()
        assert np.all(chi[~act] == 0.0)
        scale = max(1.0, float(np.abs(chi_ref).max()))
        assert float(np.abs(chi - chi_ref).max()) <= 1e-6 * scale
        sums = s.gradient_sums()
        ref = oe.gradient_sums(chi_ref, act)
        np.testing.assert_allclose(sums, ref, rtol=0, atol=1e-6 * act.sum())
        s.close()
    D, ok, infos = calculate_Deff_tensor_homogenization(ph, 1)
    assert ok and all(i["converged"] for i in infos)
    D_ref = oe.deff_tensor(ph, 1)
    np.testing.assert_allclose(D, D_ref, rtol=0, atol=DEFF_TOL)
    np.testing.assert_allclose(D, D.T, rtol=0, atol=1e-6)       # the tensor of a symmetric problem


def test_all_pore_box_gives_identity(capi):
    from openimpala_b200.effdiff import calculate_Deff_tensor_homogenization
    ph = np.ones((8, 12, 16), dtype=np.int32)
    D, ok, infos = calculate_Deff_tensor_homogenization(ph, 1)
    assert ok and all(i["iterations"] == 0 for i in infos)       # rhs = 0 -> chi = 0
    np.testing.assert_array_equal(D, np.eye(3))


def test_no_active_cells_converges_to_zero(capi):
    from openimpala_b200.effdiff import EffectiveDiffusivityHypre
    from openimpala_b200.tortuosity import Direction, SolverType
    ph = np.zeros((6, 6, 8), dtype=np.int32)
    s = EffectiveDiffusivityHypre(None, None, None, ph, 1, Direction.X, SolverType.FlexGMRES)
    assert s.solve() and s.getSolverIterations() == 0 and s.getFinalRelativeResidualNorm() == 0.0
    assert not s.getChiSolution().any()
    s.close()


def test_layers_normal_to_x(capi):
    """Pore / solid layers normal to x: chi_y = chi_z = 0 (no interface face looks along
    y or z), so D_yy = D_zz = porosity and the off-diagonal terms vanish; the x column
    equals the oracle's."""
    from openimpala_b200.effdiff import calculate_Deff_tensor_homogenization
    from oracle import oi_effdiff as oe
    ph = np.zeros((8, 8, 16), dtype=np.int32)
    ph[:, :, 4:12] = 1
    D, ok, _ = calculate_Deff_tensor_homogenization(ph, 1)
    assert ok
    assert D[1][1] == 0.5 and D[2][2] == 0.5
    assert abs(D[0][1]) + abs(D[0][2]) + abs(D[1][0]) + abs(D[2][0]) + abs(D[1][2]) + abs(D[2][1]) < 1e-12
    np.testing.assert_allclose(D, oe.deff_tensor(ph, 1), rtol=0, atol=DEFF_TOL)


def test_unsupported_solver_type_is_rejected(capi):
    from openimpala_b200.effdiff import EffectiveDiffusivityHypre
    from openimpala_b200.tortuosity import Direction, SolverType
    with pytest.raises(ValueError):
        EffectiveDiffusivityHypre(None, None, None, np.ones((4, 4, 4), np.int32), 1, Direction.X, SolverType.PCG)


def test_sample_image_tensor_golden(capi, sample_phase):
    from openimpala_b200.effdiff import calculate_Deff_tensor_homogenization
    gold = json.load(open(os.path.join(GOLDEN, "effdiff_golden.json")))
    for phase_id in (1, 0):
        D, ok, infos = calculate_Deff_tensor_homogenization(sample_phase, phase_id)
        assert ok
        np.testing.assert_allclose(D, np.array(gold[f"phase{phase_id}"]["deff"]), rtol=0, atol=DEFF_TOL)
        assert max(i["iterations"] for i in infos) < 60


def test_preconditioner_is_symmetric_cell_problem(capi):
    ph = _blobs((16, 20, 24), 21, 0.5)
    act = ph == 1
    rng = np.random.default_rng(5)
    with capi.Solver(ph.shape, 0, 1, problem=capi.OI_PROBLEM_CELL) as s:
        s.set_phase(ph)
        s.build_mask()
        u = np.where(act, rng.standard_normal(ph.shape), 0.0)
        v = np.where(act, rng.standard_normal(ph.shape), 0.0)
        mu, mv = s.apply_precond(u), s.apply_precond(v)
        a, b = float((v * mu).sum()), float((u * mv).sum())
        assert abs(a - b) <= 1e-5 * max(abs(a), abs(b))          # fp32 V-cycle
        assert float((u * mu).sum()) > 0.0
