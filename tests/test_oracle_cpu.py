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
