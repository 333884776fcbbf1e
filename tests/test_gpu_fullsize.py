"""Full-size checks (-m gpu) at BASELINE.json's configs[2] (512^3 sphere packing, tau
in Z): the oracle cannot run there in seconds, so the CUDA path is held to
size-independent properties of the domain instead --
  * integer facts agree with an independent numpy / scipy.ndimage count;
  * boundary flux is conserved (the reference's own gate, TortuosityHypre.cpp:794-803);
  * tau does not depend on the Dirichlet values (linearity);
  * the percolation mask is idempotent (mask of the mask is the mask);
  * the three stencil variants (shared-memory ring, register z-march, gather) agree;
  * the solution obeys the discrete maximum principle (vlo <= phi <= vhi).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 512
SEED, RADIUS, SOLID = 12345, 12, 0.60


@pytest.fixture(scope="module")
def capi(built_lib):
    from openimpala_b200 import capi as c
    assert c.device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return c


@pytest.fixture(scope="module")
def packing():
    from openimpala_b200 import synth
    return synth.sphere_packing(N, SEED, RADIUS, SOLID)


def _tau(capi, ph, vlo, vhi, variant=0, direction=2):
    from openimpala_b200.tortuosity import tau_from_fluxes
    with capi.Solver(ph.shape, direction, 1, vlo, vhi, stencil_variant=variant) as s:
        s.set_phase(ph)
        pc, tc = s.volume_fraction()
        n_active = s.build_mask()
        info = s.solve()
        fin, fout, ni, no = s.fluxes()
        n = ph.shape[2 - direction]
        tau, deff, conserved = tau_from_fluxes(fin, fout, n_active / ph.size, float(n), float(ph.size // n), vlo, vhi)
        return dict(pc=pc, tc=tc, n_active=n_active, info=info, fin=fin, fout=fout, tau=tau, conserved=conserved)


def test_full_size_properties(capi, packing):
    from scipy import ndimage
    ph = packing
    r = _tau(capi, ph, -1.0, 1.0)
    # integers: phase count, and the percolating count against scipy's labelling
    assert (r["pc"], r["tc"]) == (int((ph == 1).sum()), ph.size)
    lab, _ = ndimage.label(ph == 1)
    both = np.intersect1d(np.unique(lab[0][lab[0] > 0]), np.unique(lab[-1][lab[-1] > 0]))
    assert r["n_active"] == int(np.isin(lab, both).sum())
    # solve: converged by the reference's rule, flux conserved well inside its 1e-6 gate
    assert r["info"].converged and r["info"].rel_residual <= 1e-9 and r["info"].iterations <= 60
    assert r["conserved"] and abs(abs(r["fin"]) - abs(r["fout"])) <= 5e-7 * abs(r["fin"])
    assert np.isfinite(r["tau"]) and r["tau"] > 1.0
    # linearity: other Dirichlet values, same tau (1e-6, the north-star tolerance)
    r2 = _tau(capi, ph, 0.0, 3.5)
    assert abs(r2["tau"] - r["tau"]) <= 1e-6 * r["tau"]
    assert r2["n_active"] == r["n_active"]


def test_mask_idempotent_and_maximum_principle(capi, packing):
    ph = packing
    with capi.Solver(ph.shape, 2, 1, -1.0, 1.0) as s:
        s.set_phase(ph)
        n1 = s.build_mask()
        mask = s.mask()
        s.solve()
        x = s.solution()
        act = mask.astype(bool)
        assert x[act].min() >= -1.0 - 1e-6 and x[act].max() <= 1.0 + 1e-6     # discrete maximum principle (to solver accuracy)
        assert not x[~act].any()                                               # zero off the active cells
        s.set_phase(mask)                                                      # the mask as a phase field
        assert s.build_mask() == n1
        assert np.array_equal(s.mask(), mask)


def test_stencil_variants_agree_256(capi):
    from openimpala_b200 import synth
    ph = synth.sphere_packing(256, SEED, RADIUS, SOLID)
    taus = [_tau(capi, ph, -1.0, 1.0, variant=v)["tau"] for v in (0, 2, 1)]
    assert max(taus) - min(taus) <= 1e-7 * taus[0]


def test_default_large_box_path_against_independent_kernels_384(capi):
    """Above 2^25 cells the default path changes (iterations on the stream instead of CUDA graphs, the
    two-sweeps-per-pass smoother, 64-bit offsets): pin it at 384^3 against a path that shares none of its solve
    kernels -- Jacobi-preconditioned CG on the gather stencil (precond=JACOBI, stencil_variant=1) -- to 1e-7."""
    from openimpala_b200 import synth
    from openimpala_b200.tortuosity import tau_from_fluxes
    ph = synth.sphere_packing(384, SEED, RADIUS, SOLID)
    out = []
    for kw in (dict(), dict(precond=capi.OI_PRECOND_JACOBI, stencil_variant=1, maxiter=20000)):
        with capi.Solver(ph.shape, 2, 1, -1.0, 1.0, eps=1e-10, **kw) as s:
            s.set_phase(ph)
            n_active = s.build_mask()
            info = s.solve()
            fin, fout, ni, no = s.fluxes()
            tau, _, conserved = tau_from_fluxes(fin, fout, n_active / ph.size, 384.0, 384.0 * 384.0, -1.0, 1.0)
            replays, _ = s.graph_info()
            out.append(dict(tau=tau, n_active=n_active, ni=ni, no=no, info=info, replays=replays, conserved=conserved))
    a, b = out
    assert a["replays"] == 0                                   # the default path at this size is the stream path
    assert a["info"].converged and b["info"].converged and a["conserved"] and b["conserved"]
    assert (a["n_active"], a["ni"], a["no"]) == (b["n_active"], b["ni"], b["no"])
    assert a["info"].iterations <= 40 and b["info"].iterations > 10 * a["info"].iterations
    assert abs(a["tau"] - b["tau"]) <= 1e-7 * b["tau"], (a["tau"], b["tau"])


def test_sphere_packing_golden_512(capi, packing):
    """BASELINE configs[2] against the C oracle's solve of the same 512^3 image (tests/golden/packing_golden_512.json,
    Jacobi-PCG on the stored matrix to 1e-11, made by make_packing_golden_512.py): integers exact, tau / fluxes to 1e-6."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, "packing_golden_512.json")
    if not os.path.exists(path):
        pytest.skip("packing_golden_512.json not generated yet")
    g = json.load(open(path))["cases"][0]
    ph = packing
    assert hashlib.sha256(ph.tobytes()).hexdigest() == g["sha256"]
    r = _tau(capi, ph, -1.0, 1.0)
    assert r["pc"] == g["phase_count"] and r["n_active"] == g["n_active"]
    assert r["info"].converged
    assert abs(r["tau"] - g["tau"]) <= 1e-6 * g["tau"], (r["tau"], g["tau"])
    assert abs(r["fin"] - g["flux_in"]) <= 1e-6 * abs(g["flux_in"])
    assert abs(r["fout"] - g["flux_out"]) <= 1e-6 * abs(g["flux_out"])


@pytest.mark.parametrize("shape", [(256, 256, 256), (272, 296, 328)])
def test_coarse_two_sweep_pass_matches_single_sweeps(capi, shape, monkeypatch):
    """Level 1 (>= 2^21 cells, single slab) can smooth two sweeps per pass (coarse_pair_kernel, oi_coarse.cu;
    OI_COARSE_PAIR=1, opt-in because it measured slower than the single sweeps it replaces).  Same V-cycle output up to fp32 summation order, same iterations, same fluxes.  The second
    shape gives level 1 partial tiles in x and y and a partial last z-chunk."""
    from openimpala_b200 import synth
    ph = synth.sphere_packing_slab(shape, SEED, RADIUS, SOLID)
    res = {}
    for pair in ("0", "1"):
        monkeypatch.setenv("OI_COARSE_PAIR", pair)
        with capi.Solver(shape, 2, 1, -1.0, 1.0) as s:
            s.set_phase(ph)
            assert s.build_mask() > 0
            act = s.mask().astype(bool)
            r = np.where(act, np.random.default_rng(5).standard_normal(shape), 0.0)
            z = s.apply_precond(r)
            info = s.solve()
            res[pair] = (z, info.iterations, s.fluxes()[:2], s.launch_count())
    z0, it0, fl0, l0 = res["0"]
    z1, it1, fl1, l1 = res["1"]
    assert l1 < l0                                               # fewer launches: the pairs really ran
    assert float(np.abs(z0 - z1).max()) <= 2e-5 * float(np.abs(z0).max())
    assert abs(it0 - it1) <= 1
    assert abs(fl0[0] - fl1[0]) <= 1e-7 * abs(fl0[0]) and abs(fl0[1] - fl1[1]) <= 1e-7 * abs(fl0[1])
