"""CPU-side tests (run without a GPU): the oracle against the reference's own
fixtures / invariants and the committed golden vectors, the two restatements
(numpy and C) against each other, and the analytic known answers."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, SAMPLE_TIFF


@pytest.fixture(scope="module")
def o():
    from oracle import oi_numpy
    return oi_numpy


@pytest.fixture(scope="module")
def oc():
    from oracle import oi_c
    oi_c.load()
    return oi_c


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLDEN, "sample_golden.json")))


def _blobs(shape, seed, porosity=0.5, sigma=1.5):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    f = ndimage.gaussian_filter(rng.standard_normal(shape), sigma)
    return (f > np.quantile(f, 1.0 - porosity)).astype(np.int32)


# ---- reader facts pinned by the reference's tTiffReader (src/io/tTiffReader.cpp:98-219)
def test_sample_tiff_reader_facts(o, gold):
    assert hashlib.md5(open(SAMPLE_TIFF, "rb").read()).hexdigest() == gold["md5"]
    raw = o.read_tiff_raw(SAMPLE_TIFF)
    assert raw.shape == (100, 100, 100)                  # width/height/depth 100
    bo, ifds = o._read_ifds(open(SAMPLE_TIFF, "rb").read())
    assert ifds[0].get(258, (1,))[0] == 1                # BitsPerSample 1
    assert ifds[0].get(339, (1,))[0] == 1                # SampleFormat UINT
    assert ifds[0].get(277, (1,))[0] == 1                # SamplesPerPixel 1
    ph = o.threshold(raw, 0.5)
    assert ph.min() == 0 and ph.max() == 1               # thresholded min/max


def test_sample_tiff_matches_pillow(o):
    Image = pytest.importorskip("PIL.Image")
    im = Image.open(SAMPLE_TIFF)
    frames = []
    for k in range(im.n_frames):
        im.seek(k)
        frames.append(np.array(im.convert("L")) > 0)
    assert np.array_equal(np.stack(frames), o.read_tiff_raw(SAMPLE_TIFF) > 0.5)


# ---- VolumeFraction (tVolumeFraction.cpp: counts == independent loop; VF0+VF1 == 1)
def test_volume_fraction_sample(o, oc, sample_phase, gold):
    for pid in (0, 1):
        pc, tc = o.volume_fraction_counts(sample_phase, pid)
        assert pc == gold["phase_count"][str(pid)] and tc == 1_000_000
        assert oc.count_phase(sample_phase, pid) == pc
    assert gold["phase_count"]["1"] == 398309 and gold["phase_count"]["0"] == 601691


# ---- percolation mask
def test_mask_sample_golden(o, oc, sample_phase, gold):
    for case in gold["cases"]:
        m = o.activity_mask(sample_phase, case["phase"], case["direction"])
        assert int(m.sum()) == case["n_active"]
        assert hashlib.sha256(m.astype("u1").tobytes()).hexdigest() == case["mask_sha256"]
    # the reference's literal capped sweep reaches the same fixed point on the sample
    mc, n = oc.activity_mask(sample_phase, 1, 0, capped=True)
    assert n == 397743 and np.array_equal(mc.astype(bool), o.activity_mask(sample_phase, 1, 0))


@pytest.mark.parametrize("shape,seed", [((9, 11, 13), 1), ((16, 16, 16), 2), ((5, 30, 7), 3)])
def test_mask_three_ways(o, oc, shape, seed):
    ph = _blobs(shape, seed, 0.55)
    for d in range(3):
        for pid in (0, 1):
            a = o.activity_mask(ph, pid, d)
            b, _ = o.activity_mask_flood(ph, pid, d)
            c, n = oc.activity_mask(ph, pid, d)
            assert np.array_equal(a, b) and np.array_equal(a, c.astype(bool)) and n == int(a.sum())


# ---- tortuosity_fillmtx + the reference's own invariants (checkMatrixProperties)
@pytest.mark.parametrize("shape,seed", [((9, 11, 13), 1), ((16, 16, 16), 2), ((12, 7, 20), 4)])
def test_fillmtx_numpy_vs_c_and_invariants(o, oc, shape, seed):
    ph = _blobs(shape, seed, 0.6)
    for d in range(3):
        mask = o.activity_mask(ph, 1, d)
        a, rhs, x0 = o.fill_matrix(ph, mask, 1, d, -1.0, 1.0, dx=(1.0, 0.5, 2.0))
        a2, rhs2, x02 = oc.fill_matrix(ph, mask, 1, d, -1.0, 1.0, dx=(1.0, 0.5, 2.0))
        assert np.array_equal(a, a2) and np.array_equal(rhs, rhs2)
        np.testing.assert_allclose(x0, x02, rtol=0, atol=1e-15)
        assert o.check_matrix_properties(a, rhs, mask, d, -1.0, 1.0, shape)


def test_sample_matrix_invariants(o, sample_phase):
    # tTortuosity.inputs: phase 0, direction X
    mask = o.activity_mask(sample_phase, 0, 0)
    a, rhs, _ = o.fill_matrix(sample_phase, mask, 0, 0, 0.0, 1.0)
    assert o.check_matrix_properties(a, rhs, mask, 0, 0.0, 1.0, sample_phase.shape)


# ---- analytic known answers
@pytest.mark.parametrize("n", [8, 16])
def test_uniform_block(o, oc, n):
    ph = np.ones((n, n, n), dtype=np.int32)
    for d in range(3):
        assert abs(o.tortuosity(ph, 1, d).tau - (n - 1) / n) < 1e-12
        assert abs(oc.tortuosity(ph, 1, d)["tau"] - (n - 1) / n) < 1e-12


def test_half_slab_and_blocked(o, oc):
    n = 8
    ph = np.zeros((n, n, n), dtype=np.int32)
    ph[:, : n // 2, :] = 1
    r = o.tortuosity(ph, 1, 0)
    assert r.active_vf == 0.5 and abs(r.tau - (n - 1) / n) < 1e-12
    r = o.tortuosity(ph, 1, 1)
    assert r.active_vf == 0.0 and math.isnan(r.tau)
    c = oc.tortuosity(ph, 1, 1)
    assert c["n_active"] == 0 and math.isnan(c["tau"])


def test_tau_edge_conventions(o):
    shape = (4, 4, 4)
    assert math.isnan(o.tau_from_fluxes(-1.0, -1.0, 0.0, shape, 0, 0.0, 1.0)[0])          # active_vf 0
    assert math.isnan(o.tau_from_fluxes(-1.0, -1.1, 0.5, shape, 0, 0.0, 1.0)[0])          # not conserved
    assert math.isinf(o.tau_from_fluxes(0.0, 0.0, 0.5, shape, 0, 0.0, 1.0)[0])            # zero flux
    assert math.isnan(o.tau_from_fluxes(-1.0, -1.0, 0.5, shape, 0, 0.0, 1.0, converged=False)[0])


# ---- solve: C restatement against the golden (scipy) values on the sample image
@pytest.mark.parametrize("idx", [3])          # phase 1, direction X (BASELINE configs[0])
def test_sample_tau_c_oracle_vs_golden(oc, sample_phase, gold, idx):
    case = gold["cases"][idx]
    r = oc.tortuosity(sample_phase, case["phase"], case["direction"], gold["vlo"], gold["vhi"], eps=1e-10)
    assert r["n_active"] == case["n_active"]
    assert abs(r["tau"] - case["tau"]) <= 1e-7 * case["tau"]
    assert abs(case["tau"] - 3.1330740847) < 1e-9


@pytest.mark.parametrize("shape,seed", [((14, 15, 16), 5), ((20, 10, 12), 6)])
def test_tau_numpy_vs_c(o, oc, shape, seed):
    ph = _blobs(shape, seed, 0.6)
    for d in range(3):
        r = o.tortuosity(ph, 1, d, -1.0, 1.0, eps=1e-12)
        c = oc.tortuosity(ph, 1, d, -1.0, 1.0, eps=1e-12)
        assert c["n_active"] == r.n_active
        if math.isnan(r.tau):
            assert math.isnan(c["tau"])
        else:
            assert abs(c["tau"] - r.tau) <= 1e-9 * abs(r.tau)


# ---- the reference's own Krylov formulation: FlexGMRES(20) on the un-eliminated system
def test_flexgmres_restatement_known_answers(o):
    """The Krylov routine alone: a non-symmetric, diagonally dominant system against a
    direct solve, with and without restarts, exact and inexact right preconditioner,
    and HYPRE's b = 0 rule (relative to ||r0||)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(3)
    n = 300
    A = sp.random(n, n, 0.03, random_state=4, format="csr") + sp.diags(4.0 + rng.random(n))
    b = rng.standard_normal(n)
    exact = spla.spsolve(A.tocsc(), b)
    d = A.diagonal()
    for k_dim, pre in ((20, lambda v: v / d), (5, lambda v: v / d), (20, lambda v: v.copy())):
        x, it, relres, conv = o.flexgmres(A, b, np.zeros(n), pre, tol=1e-11, maxiter=400, k_dim=k_dim)
        assert conv and relres <= 1e-11
        assert np.linalg.norm(b - A @ x) <= 1.0000001e-11 * np.linalg.norm(b)
        assert np.allclose(x, exact, rtol=0, atol=1e-9)
    lu = spla.splu(A.tocsc())
    x, it, relres, conv = o.flexgmres(A, b, np.zeros(n), lu.solve, tol=1e-9, maxiter=200)
    assert conv and it == 1                                   # exact M^-1: one step
    x, it, relres, conv = o.flexgmres(A, b, np.zeros(n), lambda v: v / d, tol=1e-14, maxiter=7)
    assert not conv and it == 7                               # iteration cap reported, not hidden
    x, it, relres, conv = o.flexgmres(A, np.zeros(n), exact, lambda v: v / d, tol=1e-6, maxiter=200)
    assert conv and np.linalg.norm(A @ x) <= 1e-6 * np.linalg.norm(A @ exact)
    x, it, relres, conv = o.flexgmres(A, b, exact, lambda v: v / d, tol=1e-9, maxiter=200)
    assert conv and it == 0                                   # x0 already inside the tolerance


@pytest.mark.parametrize("shape,seed", [((20, 22, 24), 11), ((24, 28, 26), 12)])
def test_reference_formulation_flexgmres_vs_eliminated_pcg(o, shape, seed):
    """a-7 as the reference poses it (identity rows kept, non-symmetric matrix, ramp
    initial guess, FlexGMRES(20), eps 1e-9, maxiter 200, converged <=> relres <= eps)
    against the eliminated SPD solve at 1e-12: tau agrees far inside the 1e-6 bar, the
    residual of the FULL system meets the rule, and Dirichlet / inactive rows hold
    their values exactly."""
    ph = _blobs(shape, seed, 0.6)
    for d in range(3):
        ref = o.tortuosity(ph, 1, d, -1.0, 1.0, eps=1e-12)
        if math.isnan(ref.tau):
            continue
        x, mask, info = o.solve_full_flexgmres(ph, 1, d, -1.0, 1.0, eps=1e-9, maxiter=200)
        assert info["converged"] and info["iters"] <= 200 and 0.0 <= info["relres"] <= 1e-9
        fin, fout, _, _ = o.global_fluxes(x, mask, d)
        tau, _ = o.tau_from_fluxes(fin, fout, ref.active_vf, ph.shape, d, -1.0, 1.0)
        assert abs(tau - ref.tau) <= 1e-7 * ref.tau
        a, rhs, x0 = o.fill_matrix(ph, mask, 1, d, -1.0, 1.0)
        A = o.assemble_csr(a, ph.shape)
        assert np.linalg.norm(rhs - A @ x.ravel()) <= 1.0000001e-9 * np.linalg.norm(rhs)
        ident = (a[:, 0] == 1.0) & (np.abs(a[:, 1:]).sum(axis=1) == 0.0)
        assert np.array_equal(x.ravel()[ident], rhs[ident])


def test_sample_flexgmres_golden_vs_pcg_golden(gold):
    """Committed fixture (make_flexgmres_golden.py): the sample image under the reference's
    own solver settings; every case converged inside 200 iterations and its tau is within
    1e-7 of the eps-1e-12 golden -- the 1e-6 bar holds whichever Krylov method is used."""
    fg = json.load(open(os.path.join(GOLDEN, "sample_flexgmres_golden.json")))
    assert fg["eps"] == 1e-9 and fg["maxiter"] == 200 and fg["k_dim"] == 20
    by_key = {(c["phase"], c["direction"]): c for c in gold["cases"]}
    assert [(c["phase"], c["direction"]) for c in fg["cases"]] == [(1, 0), (1, 1), (1, 2)]
    for c in fg["cases"]:
        g = by_key[(c["phase"], c["direction"])]
        assert c["converged"] and c["iters"] <= 200 and c["relres"] <= 1e-9
        assert c["n_active"] == g["n_active"]
        assert abs(c["tau"] - g["tau"]) <= 1e-7 * g["tau"]
        avg = 0.5 * (abs(c["flux_in"]) + abs(c["flux_out"]))
        assert abs(abs(c["flux_in"]) - abs(c["flux_out"])) / avg <= 1e-6      # TortuosityHypre.cpp:794-803


def test_packing_golden_fixture(o, oc):
    """tests/golden/packing_golden.json (the BASELINE generator at 96^3 .. 256^3, C restatement):
    the smallest case is re-derived here by BOTH restatements, and every case keeps the
    reference's flux-conservation gate."""
    import hashlib
    from openimpala_b200 import synth
    g = json.load(open(os.path.join(GOLDEN, "packing_golden.json")))
    assert [c["n"] for c in g["cases"]] == [96, 128, 192, 256]
    c0 = g["cases"][0]
    ph = synth.sphere_packing(96, 12345, 12, 0.60)
    assert hashlib.sha256(ph.tobytes()).hexdigest() == c0["sha256"]
    assert int((ph == 1).sum()) == c0["phase_count"]
    rn = o.tortuosity(ph.astype(np.int32), 1, 2, -1.0, 1.0, eps=1e-11)
    rc = oc.tortuosity(ph.astype(np.int32), 1, 2, -1.0, 1.0, eps=1e-11)
    assert rn.n_active == rc["n_active"] == c0["n_active"]
    assert abs(rn.tau - c0["tau"]) <= 1e-9 * c0["tau"] and abs(rc["tau"] - c0["tau"]) <= 1e-9 * c0["tau"]
    for c in g["cases"]:
        avg = 0.5 * (abs(c["flux_in"]) + abs(c["flux_out"]))
        assert abs(abs(c["flux_in"]) - abs(c["flux_out"])) / avg <= 1e-8
        assert c["oracle_relres"] <= 1e-11 and 1.0 < c["tau"] < 3.0


def test_vlo_vhi_independence(oc):
    ph = _blobs((12, 12, 12), 9, 0.65)
    a = oc.tortuosity(ph, 1, 0, -1.0, 1.0, eps=1e-12)["tau"]
    b = oc.tortuosity(ph, 1, 0, 0.0, 1.0, eps=1e-12)["tau"]
    assert abs(a - b) <= 1e-9 * abs(a)


# ------------------------------------------------------------------ homogenisation cell problem (row f-1)
def test_effdiff_oracle_rows_are_symmetric_and_dominant():
    from oracle import oi_effdiff as oe
    rng = np.random.default_rng(3)
    ph = (rng.random((6, 7, 9)) < 0.6).astype(np.int32)
    for direction in range(3):
        a, rhs, x0 = oe.fill_matrix(ph, 1, direction, dx=(0.5, 1.0, 2.0))
        A = oe.assemble_csr_periodic(a, ph.shape)
        assert abs(A - A.T).max() == 0.0                          # couplings are mutual
        act = (ph == 1).ravel()
        full_diag = 2.0 * (1 / 0.25 + 1.0 + 0.25)
        assert np.all(a[act, 0] == full_diag) and np.all(a[~act, 0] == 1.0)    # F90:153-220, 124-129
        assert np.all(a[~act, 1:] == 0.0) and np.all(rhs[~act] == 0.0) and not x0.any()
        assert np.all(a[act, 0] + a[act, 1:].sum(axis=1) >= -1e-14)            # weak diagonal dominance
        # rhs closed form: (a_p - a_m) / (2 h) along the corrector direction
        h = (0.5, 1.0, 2.0)[direction]
        ax = 2 - direction
        a_p, a_m = np.roll(ph == 1, -1, axis=ax), np.roll(ph == 1, 1, axis=ax)
        closed = np.where(ph == 1, (a_p.astype(float) - a_m.astype(float)) / (2 * h), 0.0)
        np.testing.assert_allclose(rhs.reshape(ph.shape), closed, rtol=0, atol=1e-15)


def test_effdiff_oracle_analytic_cases():
    from oracle import oi_effdiff as oe
    np.testing.assert_array_equal(oe.deff_tensor(np.ones((4, 5, 6), np.int32), 1), np.eye(3))
    ph = np.zeros((4, 4, 12), dtype=np.int32)
    ph[:, :, 3:9] = 1
    D = oe.deff_tensor(ph, 1)
    assert D[1][1] == 0.5 and D[2][2] == 0.5
    assert abs(D - np.diag(np.diag(D))).max() < 1e-14
    assert np.all(oe.deff_tensor(np.zeros((3, 3, 3), np.int32), 1) == 0.0)


def test_effdiff_golden_is_symmetric_and_near_porosity(sample_phase):
    import json
    g = json.load(open(os.path.join(GOLDEN, "effdiff_golden.json")))
    for pid, n in ((1, 398309), (0, 601691)):
        D = np.array(g[f"phase{pid}"]["deff"])
        assert g[f"phase{pid}"]["n_active"] == n == int((sample_phase == pid).sum())
        np.testing.assert_allclose(D, D.T, rtol=0, atol=1e-9)
        assert np.all(np.abs(np.diag(D) - n / 1e6) < 0.02)


def test_effdiff_two_restatements_agree(oc, sample_phase):
    """numpy/scipy (oracle/oi_effdiff.py) against plain C (oracle/oi_oracle.c): rows bit-exact,
    tensors to 1e-9; the C one also reproduces the committed golden tensor of the sample image."""
    import json
    from oracle import oi_effdiff as oe
    rng = np.random.default_rng(41)
    ph = (rng.random((9, 12, 14)) < 0.55).astype(np.int32)
    for dx in ((1.0, 1.0, 1.0), (0.5, 1.25, 2.0)):
        for k in range(3):
            a1, r1, x1 = oe.fill_matrix(ph, 1, k, dx)
            a2, r2, x2 = oc.effdiff_fill_matrix(ph, 1, k, dx)
            assert np.array_equal(a1, a2) and np.array_equal(r1, r2) and not x2.any()
        D1 = oe.deff_tensor(ph, 1, dx)
        D2, iters = oc.effdiff_deff_tensor(ph, 1, dx)
        assert np.abs(D1 - D2).max() < 1e-9 and max(iters) > 0
    gold = json.load(open(os.path.join(GOLDEN, "effdiff_golden.json")))
    D, _ = oc.effdiff_deff_tensor(sample_phase, 1, eps=1e-11)
    assert np.abs(D - np.array(gold["phase1"]["deff"])).max() < 1e-8


def test_mgpcg_port_agrees_with_jacobi_pcg_goldens():
    """Third independent solve of the same systems: the CPU port of the GPU arm's MG-PCG (oracle/oi_oracle.c
    oo_solve_mgpcg; also bench.py's CPU baseline) against the Jacobi-PCG goldens of the sphere packings
    (tests/golden/packing_golden.json) -- integers exact, tau to 1e-7 (the goldens are converged to 1e-11, the
    port stops at the reference's 1e-9)."""
    import json
    from oracle import oi_c
    from openimpala_b200 import synth
    gold = json.load(open(os.path.join(GOLDEN, "packing_golden.json")))
    for case in gold["cases"][:2]:                       # 96^3, 128^3
        ph = synth.sphere_packing(case["n"], 12345, 12, 0.60).astype(np.int32)
        r = oi_c.tortuosity_mg(ph, 1, 2, -1.0, 1.0, eps=1e-9)
        assert r["n_active"] == case["n_active"]
        assert r["relres"] <= 1e-9 and r["iters"] <= 30
        assert abs(r["tau"] - case["tau"]) <= 1e-7 * case["tau"], (r["tau"], case["tau"])
    # analytic: open column, tau = (N - 1) / N
    r = oi_c.tortuosity_mg(np.ones((16, 16, 16), dtype=np.int32), 1, 0, 0.0, 1.0, eps=1e-12)
    assert abs(r["tau"] - 15.0 / 16.0) <= 1e-10
