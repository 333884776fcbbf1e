"""GPU parity tests: the CUDA path, called through the C-ABI, against the
oracle (oracle/oi_numpy.py) on the same inputs.  Integer work bit-exact, fp64
within the tolerance written beside each assert (north_star: tau within 1e-6
relative of the converged reference solve).

Modelled on the reference's own drivers: src/props/tTortuosity.cpp (construct,
checkMatrixProperties, value() finite), src/props/tVolumeFraction.cpp (counts ==
independent loop) plus the analytic cases of SURVEY 8c.
"""
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TAU_RTOL = 1e-6   # north_star tolerance on D_eff / tau


def _blobs(shape, seed, porosity=0.5, sigma=1.5):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    f = ndimage.gaussian_filter(rng.standard_normal(shape), sigma)
    return (f > np.quantile(f, 1.0 - porosity)).astype(np.int32)


@pytest.fixture(scope="module")
def capi(built_lib):
    from openimpala_b200 import capi as c
    assert c.device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return c


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("dtype", [np.uint8, np.int32])
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4097, 1_000_003])
def test_count_phase_bit_exact(capi, dtype, n):
    rng = np.random.default_rng(n + 7)
    f = rng.integers(0, 3, size=n).astype(dtype)
    for phase in (0, 1, 2, 5):
        pc, tc = capi.count_phase(f, phase)
        assert pc == int(np.count_nonzero(f == phase))
        assert tc == n


def test_volume_fraction_sample(capi, sample_phase):
    from openimpala_b200.tortuosity import VolumeFraction
    # integer facts of SURVEY 8c-3
    assert VolumeFraction(sample_phase, 1).value() == (398309, 1000000)
    assert VolumeFraction(sample_phase, 0).value() == (601691, 1000000)
    assert VolumeFraction(sample_phase.astype(np.uint8), 1).value() == (398309, 1000000)
    vf0 = VolumeFraction(sample_phase, 0).value_vf()
    vf1 = VolumeFraction(sample_phase, 1).value_vf()
    assert abs(vf0 + vf1 - 1.0) < 1e-15          # tVolumeFraction.cpp: VF0 + VF1 ~ 1


# ------------------------------------------------------------------ K2 / K7: mask + rows
CASES = [((13, 17, 20), 3, 0.55), ((24, 9, 31), 4, 0.5), ((8, 8, 8), 5, 0.7), ((33, 66, 5), 6, 0.6),
         ((40, 40, 40), 7, 0.45), ((7, 130, 70), 8, 0.5)]


@pytest.mark.parametrize("shape,seed,por", CASES)
@pytest.mark.parametrize("direction", [0, 1, 2])
def test_mask_rows_and_operator(capi, shape, seed, por, direction):
    from oracle import oi_numpy as o
    ph = _blobs(shape, seed, por)
    for phase_id in (1, 0):
        mask = o.activity_mask(ph, phase_id, direction)
        with capi.Solver(shape, direction, phase_id, vlo=-1.0, vhi=1.0) as s:
            s.set_phase(ph)
            assert s.volume_fraction() == o.volume_fraction_counts(ph, phase_id)
            n_active = s.build_mask()
            assert n_active == int(mask.sum())                       # bit-exact count
            if n_active == 0:
                continue
            assert np.array_equal(s.mask().astype(bool), mask)       # bit-exact mask
            a, rhs, x0 = o.fill_matrix(ph, mask, phase_id, direction, -1.0, 1.0)
            assert np.array_equal(s.matrix_rows().reshape(-1, 7), a)  # coefficients are exact
            assert np.array_equal(s.rhs().ravel(), rhs)
            np.testing.assert_allclose(s.initial_guess().ravel(), x0, rtol=0, atol=1e-15)
            assert s.check_matrix_properties()
            assert o.check_matrix_properties(s.matrix_rows(), s.rhs(), mask, direction, -1.0, 1.0, shape)
            # K3: y = A_elim x against the assembled matrix
            A = o.assemble_csr(a, shape)
            Auu, bu, unk, xf = o.eliminate_dirichlet(A, rhs, x0, shape, mask, direction)
            rng = np.random.default_rng(seed)
            x = np.where(unk, rng.standard_normal(unk.size), 0.0)
            y_ref = np.zeros(unk.size)
            y_ref[unk] = Auu @ x[unk]
            y = s.apply_operator(x.reshape(shape)).ravel()
            np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("shape,seed,por", [((13, 17, 20), 3, 0.55), ((40, 40, 40), 7, 0.45), ((7, 130, 70), 8, 0.5),
                                            ((64, 48, 96), 9, 0.42)])
@pytest.mark.parametrize("direction", [0, 1, 2])
@pytest.mark.parametrize("ccl", ["0", "1"])
def test_mask_both_labelling_schedules(capi, shape, seed, por, direction, ccl, monkeypatch):
    """The percolation labelling merges either in one pass over the box (OI_CCL=0, the default below 2^25 cells)
    or slice by slice (OI_CCL=1, the default above): union-find results do not depend on the order of the
    unions, so both must give the oracle's mask bit for bit (near-threshold porosities: many components)."""
    from oracle import oi_numpy as o
    monkeypatch.setenv("OI_CCL", ccl)
    ph = _blobs(shape, seed, por)
    for phase_id in (1, 0):
        mask = o.activity_mask(ph, phase_id, direction)
        with capi.Solver(shape, direction, phase_id, vlo=-1.0, vhi=1.0) as s:
            s.set_phase(ph)
            assert s.build_mask() == int(mask.sum())
            if mask.any():
                assert np.array_equal(s.mask().astype(bool), mask)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("shape,seed,por", CASES[:5])
def test_tau_matches_oracle(capi, shape, seed, por, variant):
    from oracle import oi_numpy as o
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre
    ph = _blobs(shape, seed, por)
    for direction in (0, 1, 2):
        ref = o.tortuosity(ph, 1, direction, -1.0, 1.0, eps=1e-13)
        t = TortuosityHypre(None, None, None, ph, ref.active_vf, 1, Direction(direction),
                            SolverType.FlexGMRES, "", -1.0, 1.0, stencil_variant=variant)
        assert t.checkMatrixProperties()
        tau = t.value()
        assert t.getActiveVolumeFraction() == ref.active_vf
        if math.isnan(ref.tau):
            assert math.isnan(tau)
            continue
        assert t.getSolverConverged()
        assert t.getFinalRelativeResidualNorm() <= 1e-9
        assert abs(tau - ref.tau) <= TAU_RTOL * abs(ref.tau), (tau, ref.tau)
        np.testing.assert_allclose(t.getFluxIn(), ref.flux_in, rtol=TAU_RTOL)
        np.testing.assert_allclose(t.getFluxOut(), ref.flux_out, rtol=TAU_RTOL)
        t.close()


def test_jacobi_pcg_matches_mg(capi):
    from openimpala_b200 import capi as c
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre
    ph = _blobs((20, 24, 28), 11, 0.55)
    taus = []
    for pre in (c.OI_PRECOND_MG, c.OI_PRECOND_JACOBI):
        from openimpala_b200.tortuosity import ParmParse
        ParmParse.table = {"hypre.maxiter": 2000}
        try:
            t = TortuosityHypre(None, None, None, ph, 0.5, 1, Direction.X, SolverType.PCG, "", 0.0, 1.0,
                                precond=pre)
            taus.append(t.value())
            assert t.getSolverConverged()
        finally:
            ParmParse.table = {}
    assert abs(taus[0] - taus[1]) <= TAU_RTOL * abs(taus[0])


# ------------------------------------------------------------------ analytic known answers
@pytest.mark.parametrize("n", [8, 16, 33])
@pytest.mark.parametrize("direction", [0, 1, 2])
def test_uniform_block(capi, n, direction):
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre
    ph = np.ones((n, n, n), dtype=np.int32)
    t = TortuosityHypre(None, None, None, ph, 1.0, 1, Direction(direction), SolverType.FlexGMRES, "")
    assert t.getActiveVolumeFraction() == 1.0
    assert abs(t.value() - (n - 1) / n) <= 1e-12
    assert t.getSolverConverged()


def test_half_slab_and_blocked(capi):
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre
    n = 8
    ph = np.zeros((n, n, n), dtype=np.int32)
    ph[:, : n // 2, :] = 1                    # phase occupies y < n/2
    t = TortuosityHypre(None, None, None, ph, 0.5, 1, Direction.X, SolverType.FlexGMRES, "")
    assert t.getActiveVolumeFraction() == 0.5
    assert abs(t.value() - (n - 1) / n) <= 1e-12
    t = TortuosityHypre(None, None, None, ph, 0.5, 1, Direction.Y, SolverType.FlexGMRES, "")
    assert t.getActiveVolumeFraction() == 0.0           # no outlet seeds -> empty mask
    assert math.isnan(t.value())
    assert t.checkMatrixProperties()


def test_equal_potentials_gives_inf(capi):
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre
    ph = np.ones((6, 6, 6), dtype=np.int32)
    t = TortuosityHypre(None, None, None, ph, 1.0, 1, Direction.Z, SolverType.FlexGMRES, "", 0.5, 0.5)
    assert math.isinf(t.value())               # avg flux ~ 0 -> +Inf (TortuosityHypre.cpp:846-851)


def test_flux_gate_returns_nan_like_the_reference(capi, sample_phase):
    """Default (flux_polish = 0) is the reference's behaviour: the solve stops on the residual rule alone and value()
    returns NaN when the boundary fluxes disagree by more than 1e-6 (TortuosityHypre.cpp:794-823).  A loose
    hypre.eps makes the gate fire; the opt-in polish keeps iterating and returns a finite tau."""
    from openimpala_b200.tortuosity import Direction, ParmParse, SolverType, TortuosityHypre
    ParmParse.table["hypre.eps"] = 1e-4
    try:
        t = TortuosityHypre(None, None, None, sample_phase, 0.4, 1, Direction.X, SolverType.FlexGMRES, "", -1.0, 1.0)
        tau = t.value()
        assert t.getSolverConverged() and t.getFinalRelativeResidualNorm() <= 1e-4
        fin, fout = abs(t.getFluxIn()), abs(t.getFluxOut())
        mismatch = abs(fin - fout) / (0.5 * (fin + fout))
        assert mismatch > 1e-6, "pick a looser eps: the gate has to fire for this test to mean anything"
        assert math.isnan(tau)                                               # the reference's NaN
        t.close()
        t = TortuosityHypre(None, None, None, sample_phase, 0.4, 1, Direction.X, SolverType.FlexGMRES, "", -1.0, 1.0,
                            flux_polish=1)
        tau = t.value()
        fin, fout = abs(t.getFluxIn()), abs(t.getFluxOut())
        assert t.getSolverConverged() and math.isfinite(tau)
        assert abs(fin - fout) / (0.5 * (fin + fout)) <= 1e-6
        assert abs(tau - 3.1330740847) <= 1e-4 * 3.1330740847               # SURVEY 8c-3 (sample, phase 1, X)
        t.close()
        # a polish round that runs into hypre.maxiter must not downgrade a solve confirmed at eps
        ParmParse.table["hypre.maxiter"] = 6
        ParmParse.table["hypre.eps"] = 1e-2
        t = TortuosityHypre(None, None, None, sample_phase, 0.4, 1, Direction.X, SolverType.FlexGMRES, "", -1.0, 1.0,
                            flux_polish=1)
        t.value()
        if t.getSolverIterations() >= 6 and t.getFinalRelativeResidualNorm() <= 1e-2:
            assert t.getSolverConverged()
        t.close()
    finally:
        ParmParse.table.pop("hypre.eps", None)
        ParmParse.table.pop("hypre.maxiter", None)


def test_anisotropic_cell_size(capi):
    from oracle import oi_numpy as o
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre
    ph = _blobs((18, 20, 22), 21, 0.6)
    dx = (0.5, 1.0, 2.0)
    for d in (0, 2):
        ref = o.tortuosity(ph, 1, d, 0.0, 1.0, eps=1e-13, dx=dx)
        t = TortuosityHypre({"dx": dx}, None, None, ph, 0.6, 1, Direction(d), SolverType.FlexGMRES, "")
        assert abs(t.value() - ref.tau) <= TAU_RTOL * abs(ref.tau)


@pytest.mark.parametrize("dx,direction,max_iters", [((1.0, 1.0, 5.0), 2, 70), ((1.0, 1.0, 5.0), 0, 70),
                                                    ((3.0, 1.0, 1.0), 0, 60), ((1.0, 4.0, 1.0), 1, 70)])
def test_strongly_anisotropic_cells_semicoarsen(capi, dx, direction, max_iters):
    """Voxels five times longer along one axis (FIB-SEM stacks): the hierarchy coarsens only
    the strongly coupled axes until the couplings even out, so the iteration count stays far
    from hypre.maxiter = 200 (full 2x2x2 coarsening needed 131 at 256^3) and tau still
    matches the oracle."""
    from oracle import oi_numpy as o
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre
    ph = _blobs((40, 44, 48), 33, 0.55, sigma=2.0)
    ref = o.tortuosity(ph, 1, direction, -1.0, 1.0, eps=1e-12, dx=dx)
    t = TortuosityHypre({"dx": dx}, None, None, ph, 0.55, 1, Direction(direction), SolverType.FlexGMRES, "", -1.0, 1.0)
    tau = t.value()
    assert t.getSolverConverged() and t.getSolverIterations() <= max_iters
    assert abs(tau - ref.tau) <= TAU_RTOL * abs(ref.tau)
    big = _blobs((128, 128, 128), 3, 0.5, sigma=2.0)
    with capi.Solver(big.shape, direction, 1, -1.0, 1.0, dx=dx) as s:
        s.set_phase(big)
        s.build_mask()
        info = s.solve()
        assert info.converged and info.iterations <= max_iters


# ------------------------------------------------------------------ the reference's sample image
def test_sample_image_golden(capi, sample_phase):
    """BASELINE configs[0]/[1]: tau in X/Y/Z for both phases + VF; golden values
    from tests/golden/make_golden.py (oracle at eps 1e-12)."""
    import hashlib
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre
    gold = json.load(open(os.path.join(GOLDEN, "sample_golden.json")))
    # the same image under the reference's own formulation and solver settings (FlexGMRES(20)
    # on the un-eliminated system, eps 1e-9, maxiter 200): make_flexgmres_golden.py
    fgm = {(c["phase"], c["direction"]): c
           for c in json.load(open(os.path.join(GOLDEN, "sample_flexgmres_golden.json")))["cases"]}
    for case in gold["cases"]:
        t = TortuosityHypre(None, None, None, sample_phase, 0.4, case["phase"], Direction(case["direction"]),
                            SolverType.FlexGMRES, "", gold["vlo"], gold["vhi"])
        assert t._n_active == case["n_active"]
        assert hashlib.sha256(t.active_mask().tobytes()).hexdigest() == case["mask_sha256"]
        assert t.checkMatrixProperties()
        tau = t.value()                        # default eps 1e-9, maxiter 200
        assert t.getSolverConverged() and t.getSolverIterations() <= 200
        assert abs(tau - case["tau"]) <= TAU_RTOL * case["tau"], (case, tau)
        fg = fgm.get((case["phase"], case["direction"]))        # phase 1 only
        assert fg is None or abs(tau - fg["tau"]) <= TAU_RTOL * fg["tau"], (fg, tau)
        fi, fo, ni, no = t.solver.fluxes()
        assert (ni, no) == (case["n_in"], case["n_out"])
        t.close()


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_sphere_packing_golden(capi, idx):
    """The BASELINE workload generator at 96^3 ... 256^3 against the C restatement's answers
    (tests/golden/packing_golden.json, made by make_packing_golden.py with Jacobi-PCG at 1e-11):
    same packing (sha256), same phase and percolating counts, tau and both fluxes within 1e-6."""
    import hashlib
    from openimpala_b200 import synth
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre, VolumeFraction
    gold = json.load(open(os.path.join(GOLDEN, "packing_golden.json")))
    case = gold["cases"][idx]
    ph = synth.sphere_packing(case["n"], 12345, 12, 0.60)
    assert hashlib.sha256(ph.tobytes()).hexdigest() == case["sha256"]
    pc, tc = VolumeFraction(ph, 1).value()
    assert (pc, tc) == (case["phase_count"], case["n"] ** 3)
    t = TortuosityHypre(None, None, None, ph, pc / tc, 1, Direction.Z, SolverType.FlexGMRES, "", gold["vlo"], gold["vhi"])
    assert t._n_active == case["n_active"]
    tau = t.value()                                # default eps 1e-9, maxiter 200
    assert t.getSolverConverged()
    assert abs(tau - case["tau"]) <= TAU_RTOL * case["tau"], (case, tau)
    assert abs(t.getFluxIn() - case["flux_in"]) <= TAU_RTOL * abs(case["flux_in"])
    assert abs(t.getFluxOut() - case["flux_out"]) <= TAU_RTOL * abs(case["flux_out"])
    t.close()


def test_preconditioner_is_symmetric(capi):
    """PCG needs M = M^T: <M a, b> == <a, M b> on random vectors."""
    ph = _blobs((24, 28, 32), 31, 0.6)
    with capi.Solver(ph.shape, 2, 1) as s:
        s.set_phase(ph)
        s.build_mask()
        unk = (s.matrix_rows()[..., 0] != 1.0) | (s.matrix_rows()[..., 1:] != 0).any(axis=-1)
        rng = np.random.default_rng(5)
        a = np.where(unk, rng.standard_normal(ph.shape), 0.0)
        b = np.where(unk, rng.standard_normal(ph.shape), 0.0)
        ma, mb = s.apply_precond(a), s.apply_precond(b)
        lhs, rhs = float((ma * b).sum()), float((a * mb).sum())
        # the V-cycle runs in fp32 (mg_t): symmetric up to single-precision rounding
        assert abs(lhs - rhs) <= 5e-5 * max(abs(lhs), abs(rhs))
        assert float((ma * a).sum()) > 0.0


# ------------------------------------------------------------------ a-3: tortuosity_remspot
def _remspot_cases():
    rng = np.random.default_rng(42)
    noise = (rng.random((14, 18, 22)) < 0.5).astype(np.int32)            # many isolated voxels
    kk, jj, ii = np.indices((10, 12, 16))
    checker = ((kk + jj + ii) & 1).astype(np.int32)                       # every voxel isolated: order matters
    blobs = _blobs((20, 20, 20), 3, 0.5, sigma=0.8)
    mixed = blobs.copy()
    mixed[5:12, 5:12, 5:12] = checker[:7, :7, :7]
    return [noise, checker, blobs, mixed]


@pytest.mark.parametrize("case", range(4))
@pytest.mark.parametrize("passes", [1, 2])
def test_remspot_matches_sequential_reference_order(capi, case, passes):
    from oracle import oi_c, oi_numpy as o
    ph = _remspot_cases()[case]
    ref = oi_c.remspot(ph, passes)
    if ph.size <= 4000:
        assert np.array_equal(ref, o.remspot(ph, passes))                 # the two restatements agree
    for phase_id in (1, 0):
        with capi.Solver(ph.shape, 0, phase_id) as s:
            s.set_phase(ph)
            s.remspot(passes)
            n_active = s.build_mask()
            mask_ref = o.activity_mask(ref, phase_id, 0)
            assert n_active == int(mask_ref.sum())
            if n_active:
                assert np.array_equal(s.mask().astype(bool), mask_ref)


def test_remspot_rejects_non_binary_field(capi):
    ph = np.full((6, 6, 6), 2, dtype=np.int32)
    with capi.Solver(ph.shape, 0, 1) as s:
        s.set_phase(ph)
        with pytest.raises(capi.OiError):
            s.remspot(1)
        s.remspot(0)                                                       # 0 passes: skipped, like the reference


def test_remspot_through_class_parmparse(capi):
    from oracle import oi_c, oi_numpy as o
    from openimpala_b200.tortuosity import Direction, ParmParse, SolverType, TortuosityHypre
    rng = np.random.default_rng(7)
    ph = _blobs((18, 18, 18), 12, 0.6)
    ph[rng.random(ph.shape) < 0.03] ^= 1                                   # salt-and-pepper
    ParmParse.table = {"tortuosity.remspot_passes": 1}
    try:
        t = TortuosityHypre(None, None, None, ph, 0.6, 1, Direction.Y, SolverType.FlexGMRES, "", -1.0, 1.0)
    finally:
        ParmParse.table = {}
    ref = o.tortuosity(oi_c.remspot(ph, 1), 1, 1, -1.0, 1.0, eps=1e-13)
    assert t._n_active == ref.n_active
    assert abs(t.value() - ref.tau) <= TAU_RTOL * abs(ref.tau)


# ------------------------------------------------------------------ streamed upload (row f-2)
@pytest.mark.parametrize("chunk", [1, 5, 16, 100])
def test_streamed_upload_matches_one_shot(capi, chunk):
    from oracle import oi_numpy as o
    shape = (37, 20, 24)
    ph = _blobs(shape, 31, 0.55).astype(np.uint8)
    with capi.Solver(shape, 2, 1, -1.0, 1.0) as a, capi.Solver(shape, 2, 1, -1.0, 1.0) as b:
        a.set_phase(ph)
        calls = []

        def read(z0, nz, out):
            calls.append((z0, nz))
            out[...] = ph[z0:z0 + nz]
        b.set_phase_streamed(read, planes_per_chunk=chunk)
        assert sum(n for _, n in calls) == shape[0] and calls[0][0] == 0
        assert a.volume_fraction() == b.volume_fraction() == o.volume_fraction_counts(ph, 1)
        assert a.build_mask() == b.build_mask()
        assert np.array_equal(a.mask(), b.mask())
        ia, ib = a.solve(), b.solve()
        assert ia.iterations == ib.iterations and a.fluxes() == b.fluxes()
        # the same handle can be re-filled either way
        b.set_phase(ph)
        assert b.build_mask() == a.build_mask()


def test_streamed_upload_rejects_incomplete_slab(capi):
    import ctypes as C
    shape = (8, 4, 4)
    with capi.Solver(shape, 2, 1) as s:
        lib = capi.load()
        assert lib.oi_phase_stream_begin(s._h, 3) == 0
        buf = C.POINTER(C.c_uint8)()
        assert lib.oi_phase_stream_buffer(s._h, 0, C.byref(buf)) == 0
        assert lib.oi_phase_stream_submit(s._h, 0, 0, 3) == 0
        assert lib.oi_phase_stream_submit(s._h, 0, 6, 3) != 0          # runs past the slab
        assert lib.oi_phase_stream_submit(s._h, 0, 0, 4) != 0          # larger than the staging buffer
        assert lib.oi_phase_stream_end(s._h) != 0                      # 3 of 8 planes only
        assert b"every plane" in lib.oi_last_error()


# ------------------------------------------------------------------ two sweeps per pass
@pytest.mark.parametrize("shape,seed,por", [((40, 37, 100), 51, 0.5), ((70, 64, 64), 52, 0.45), ((33, 16, 8), 53, 0.6),
                                            ((20, 130, 132), 54, 0.55)])
@pytest.mark.parametrize("direction", [0, 2])
@pytest.mark.parametrize("pair_variant", ["1", "2", "3"])
def test_pair_kernel_matches_single_sweeps(capi, shape, seed, por, direction, pair_variant, monkeypatch):
    """The temporally blocked smoother (two sweeps per pass, oi_level0_pair.cu; OI_PAIR selects
    the variant, 0 = off) against the single-sweep ring kernels: same V-cycle output (fp32 rounding only),
    same iteration count, same tau.  Shapes cover partial tiles in x and y and z-chunk boundaries."""
    ph = _blobs(shape, seed, por)
    rng = np.random.default_rng(seed)
    res = {}
    for no_pair in ("1", "0"):
        monkeypatch.setenv("OI_PAIR", "0" if no_pair == "1" else pair_variant)
        with capi.Solver(shape, direction, 1, -1.0, 1.0) as s:
            s.set_phase(ph)
            if s.build_mask() == 0:
                pytest.skip("nothing percolates")
            act = s.mask().astype(bool)
            r = np.where(act, np.random.default_rng(seed).standard_normal(shape), 0.0)
            z = s.apply_precond(r)
            info = s.solve()
            res[no_pair] = (z, info.iterations, s.fluxes()[:2], s.launch_count())
    z1, it1, fl1, l1 = res["1"]
    z0, it0, fl0, l0 = res["0"]
    assert l0 < l1                                               # fewer launches: the pairs really ran
    scale = float(np.abs(z1).max())
    assert float(np.abs(z0 - z1).max()) <= 2e-5 * scale          # fp32 V-cycle, different summation order only
    assert abs(it0 - it1) <= 1
    assert abs(fl0[0] - fl1[0]) <= 1e-7 * abs(fl1[0]) and abs(fl0[1] - fl1[1]) <= 1e-7 * abs(fl1[1])


@pytest.mark.parametrize("shape,seed,por", [((40, 37, 96), 61, 0.5), ((70, 50, 80), 62, 0.45), ((130, 24, 16), 63, 0.6),
                                            ((20, 130, 144), 64, 0.55)])
@pytest.mark.parametrize("direction", [0, 2])
def test_tma_ring_matches_cp_async_ring(capi, shape, seed, por, direction, monkeypatch):
    """OI_TMA=1 stages the z-planes of the level-0 operator apply and smoother with cp.async.bulk.tensor + mbarriers
    (oi_level0_tma.cu) instead of per-thread cp.async.  Same tile, same arithmetic in the same order: A p bit-identical,
    the V-cycle to fp32 rounding of the parts that still differ (none expected), same iterations and fluxes.  Shapes
    have nx % 16 == 0 and cover partial tiles in x and y, z-chunk boundaries and boxes thinner than the ring."""
    ph = _blobs(shape, seed, por)
    res = {}
    for tma in ("0", "1"):
        monkeypatch.setenv("OI_TMA", tma)
        monkeypatch.setenv("OI_PAIR", "0")                       # single sweeps: the kernels under test
        with capi.Solver(shape, direction, 1, -1.0, 1.0) as s:
            s.set_phase(ph)
            if s.build_mask() == 0:
                pytest.skip("nothing percolates")
            act = s.mask().astype(bool)
            x = np.where(act, np.random.default_rng(seed).standard_normal(shape), 0.0)
            y = s.apply_operator(x)
            z = s.apply_precond(x)
            info = s.solve()
            res[tma] = (y, z, info.iterations, info.rel_residual, s.fluxes()[:2])
    y0, z0, it0, rr0, fl0 = res["0"]
    y1, z1, it1, rr1, fl1 = res["1"]
    assert np.array_equal(y0, y1)
    assert float(np.abs(z0 - z1).max()) <= 1e-6 * float(np.abs(z0).max())
    assert it0 == it1
    assert abs(fl0[0] - fl1[0]) <= 1e-9 * abs(fl0[0]) and abs(fl0[1] - fl1[1]) <= 1e-9 * abs(fl0[1])


# ------------------------------------------------------------------ iterations replayed as CUDA graphs
@pytest.mark.parametrize("case", ["sample_x", "blobs_z", "blobs_jacobi", "odd_nx"])
def test_iteration_graph_matches_stream_path(capi, sample_phase, case, monkeypatch):
    """Small boxes are launch bound, so every PCG iteration after the first is one graph launch
    (oi_graph_info; OI_GRAPH=0 keeps the stream path).  Same iteration count, same residual
    history end point, same solution and fluxes; the graphs survive a second solve on the handle
    and are re-captured after the mask is rebuilt."""
    if case == "sample_x":
        ph, direction, kw = sample_phase, 0, {}
    elif case == "blobs_z":
        ph, direction, kw = _blobs((48, 40, 64), 61, 0.55), 2, {}
    elif case == "blobs_jacobi":
        ph, direction, kw = _blobs((24, 20, 28), 62, 0.6), 1, {"precond": 1, "maxiter": 5000}
    else:
        ph, direction, kw = _blobs((30, 26, 37), 63, 0.6), 0, {}      # nx % 4 != 0: z-march kernels
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("OI_GRAPH", mode)
        with capi.Solver(ph.shape, direction, 1, -1.0, 1.0, **kw) as s:
            s.set_phase(ph)
            if s.build_mask() == 0:
                pytest.skip("nothing percolates")
            x_start = s.solution()
            info = s.solve()
            replays, nodes = s.graph_info()
            first = (info.iterations, info.rel_residual, info.converged, s.solution(), s.fluxes()[:2], replays, nodes)
            # again from the same start on the same handle: cached graphs, same answer
            s.set_solution(x_start)
            info2 = s.solve()
            assert info2.iterations == info.iterations and info2.converged == info.converged
            np.testing.assert_allclose(s.solution(), first[3], rtol=0, atol=1e-12)
            # mask rebuilt: graphs dropped and captured afresh
            s.build_mask()
            info3 = s.solve()
            assert info3.iterations == info.iterations
            np.testing.assert_allclose(s.solution(), first[3], rtol=0, atol=1e-12)
            out[mode] = first + (s.graph_info()[0],)
    it0, rel0, conv0, x0, fl0, rep0, nodes0, rep0_all = out["0"]
    it1, rel1, conv1, x1, fl1, rep1, nodes1, rep1_all = out["1"]
    assert conv0 and conv1 and it0 == it1 and it1 >= 3
    assert rep0 == 0 and nodes0 == 0 and rep0_all == 0                    # OI_GRAPH=0: stream path only
    assert rep1 == it1 - 1 and nodes1 >= 4 and rep1_all == 3 * (it1 - 1)  # every iteration but the first (Jacobi-PCG: 4 kernels)
    assert abs(rel0 - rel1) <= 1e-6 * rel0
    np.testing.assert_allclose(x1, x0, rtol=0, atol=1e-12)
    assert abs(fl0[0] - fl1[0]) <= 1e-11 * abs(fl0[0]) and abs(fl0[1] - fl1[1]) <= 1e-11 * abs(fl0[1])


def test_iteration_graph_default_and_cell_problem(capi, monkeypatch):
    """Default gating: graphs on for a small single slab, for the periodic cell problem too
    (its wrapped ghost planes are copy nodes of the graph)."""
    from openimpala_b200.effdiff import EffectiveDiffusivityHypre
    monkeypatch.delenv("OI_GRAPH", raising=False)
    ph = _blobs((32, 28, 24), 64, 0.6)
    with capi.Solver(ph.shape, 2, 1, -1.0, 1.0) as s:
        s.set_phase(ph)
        assert s.build_mask() > 0
        info = s.solve()
        assert info.converged and s.graph_info()[0] == info.iterations - 1
    chi = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("OI_GRAPH", mode)
        with capi.Solver(ph.shape, 0, 1, problem=1, maxiter=1000) as s:
            s.set_phase(ph)
            s.build_mask()
            info = s.solve()
            assert info.converged
            chi[mode] = (info.iterations, s.solution(), s.graph_info()[0])
    assert chi["0"][0] == chi["1"][0] and chi["0"][2] == 0 and chi["1"][2] == chi["1"][0] - 1
    np.testing.assert_allclose(chi["1"][1], chi["0"][1], rtol=0, atol=1e-12)


# ------------------------------------------------------------------ one-CTA coarse tail
@pytest.mark.parametrize("case", ["sample_z", "blobs_x", "whole_hierarchy", "anisotropic", "cell_problem", "odd_nx"])
def test_coarse_tail_matches_level_kernels(capi, sample_phase, case, monkeypatch):
    """OI_TAIL=1: every level of at most 4096 cells and everything below it is cycled by ONE
    kernel (oi_coarse_tail.cuh) instead of ~11 launches per level.  Same V-cycle up to fp32
    summation order: preconditioner output, iteration count, result."""
    kw, dx, problem = {}, (1.0, 1.0, 1.0), 0
    if case == "sample_z":
        ph, direction = sample_phase, 2                              # tail = 13^3, 7^3, 4^3
    elif case == "blobs_x":
        ph, direction = _blobs((48, 40, 64), 71, 0.55), 0            # tail = 12x10x16 and below
    elif case == "whole_hierarchy":
        ph, direction = _blobs((24, 20, 28), 72, 0.6), 1             # level 1 already fits: no level kernels at all
    elif case == "anisotropic":
        ph, direction, dx = _blobs((40, 36, 32), 73, 0.6), 2, (1.0, 1.0, 4.0)   # semicoarsened levels in the tail
    elif case == "cell_problem":
        ph, direction, problem = _blobs((32, 28, 24), 74, 0.6), 0, 1           # periodic: wrapped indices
        kw = {"maxiter": 1000}
    else:
        ph, direction = _blobs((30, 26, 37), 75, 0.6), 0
    out = {}
    for tail in ("0", "1"):
        monkeypatch.setenv("OI_TAIL", tail)
        with capi.Solver(ph.shape, direction, 1, -1.0, 1.0, dx=dx, problem=problem, **kw) as s:
            s.set_phase(ph)
            if s.build_mask() == 0:
                pytest.skip("nothing percolates")
            act = s.mask().astype(bool)
            r = np.where(act, np.random.default_rng(7).standard_normal(ph.shape), 0.0)
            l0 = s.launch_count()
            z = s.apply_precond(r)
            per_cycle = s.launch_count() - l0
            info = s.solve()
            assert info.converged
            out[tail] = (z, per_cycle, info.iterations, s.solution())
    z0, n0, it0, x0 = out["0"]
    z1, n1, it1, x1 = out["1"]
    assert n1 <= n0 - 15                                             # at least two levels' worth of launches gone
    scale = float(np.abs(z0).max())
    assert float(np.abs(z1 - z0).max()) <= 2e-5 * scale
    assert abs(it1 - it0) <= 1
    # two converged solves (relative residual 1e-9) of the same system, not the same iterates
    np.testing.assert_allclose(x1, x0, rtol=0, atol=1e-5 * max(1.0, float(np.abs(x0).max())))
