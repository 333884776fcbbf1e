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
