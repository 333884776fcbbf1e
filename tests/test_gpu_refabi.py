"""The reference's two existing C-ABI kernels under their own names and Fortran calling convention
(include/openimpala_b200.h: tortuosity_fillmtx, tortuosity_remspot), device implementations in
openimpala_b200/csrc/oi_refabi.cu, against the oracle restatements of
src/props/TortuosityHypreFill.F90:44-314 and src/props/Tortuosity_filcc.F90:88-177.
Everything here is integer work or exactly representable coefficients: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


def _tiles(shape, tile):
    nz, ny, nx = shape
    for k0 in range(0, nz, tile[0]):
        for j0 in range(0, ny, tile[1]):
            for i0 in range(0, nx, tile[2]):
                yield ((i0, j0, k0), (min(i0 + tile[2], nx) - 1, min(j0 + tile[1], ny) - 1, min(k0 + tile[0], nz) - 1))


@pytest.mark.parametrize("shape,seed,tile", [((13, 17, 20), 3, (13, 17, 20)), ((24, 9, 31), 4, (8, 8, 1024000)),
                                             ((16, 16, 16), 5, (5, 7, 6))])
@pytest.mark.parametrize("direction", [0, 1, 2])
@pytest.mark.parametrize("dx", [(1.0, 1.0, 1.0), (0.5, 2.0, 1.25)])
def test_tortuosity_fillmtx_matches_oracle(capi, shape, seed, tile, direction, dx):
    """Whole domain as one box and as MFIter-style tiles (AMReX's default tile is (1024000, 8, 8)), fields with one
    ghost cell as in TortuosityHypre.cpp:590-632; a, rhs, xinit bit-identical to the restated Fortran."""
    from oracle import oi_numpy as o
    ph = _blobs(shape, seed, 0.55)
    nz, ny, nx = shape
    for phase_id in (1, 0):
        mask = o.activity_mask(ph, phase_id, direction)
        a_ref, rhs_ref, x_ref = o.fill_matrix(ph, mask, phase_id, direction, -1.0, 2.5, dx)
        a_ref = a_ref.reshape(nz, ny, nx, 7); rhs_ref = rhs_ref.reshape(shape); x_ref = x_ref.reshape(shape)
        # one ghost cell: phase ghosts hold the phase id (worst case), mask ghosts are 0 (TortuosityHypre.cpp:309, 522)
        pg = np.full((nz + 2, ny + 2, nx + 2), phase_id, dtype=np.int32)
        pg[1:-1, 1:-1, 1:-1] = ph
        mg = np.zeros((nz + 2, ny + 2, nx + 2), dtype=np.int32)
        mg[1:-1, 1:-1, 1:-1] = mask
        dxinv = [1.0 / d ** 2 for d in dx]
        for lo, hi in _tiles(shape, tile):
            a, rhs, x = capi.ref_tortuosity_fillmtx(pg, (-1, -1, -1), mg, (-1, -1, -1), lo, hi, (0, 0, 0),
                                                    (nx - 1, ny - 1, nz - 1), dxinv, -1.0, 2.5, phase_id, direction)
            sl = (slice(lo[2], hi[2] + 1), slice(lo[1], hi[1] + 1), slice(lo[0], hi[0] + 1))
            assert np.array_equal(a, a_ref[sl].reshape(-1, 7))
            assert np.array_equal(rhs, rhs_ref[sl].ravel())
            assert np.array_equal(x, x_ref[sl].ravel())


def _remspot_box_numpy(q, q_lo, lo, hi, domlo, domhi):
    """Tortuosity_filcc.F90:88-177 literally, on a box of an array with its own lower bound."""
    q = q.copy()
    for k in range(lo[2], hi[2] + 1):
        for j in range(lo[1], hi[1] + 1):
            for i in range(lo[0], hi[0] + 1):
                c = q[k - q_lo[2], j - q_lo[1], i - q_lo[0]]
                same = False
                for d, (di, dj, dk) in enumerate(((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1))):
                    ax, side = d // 2, d % 2
                    if (i, j, k)[ax] == (domhi if side else domlo)[ax]:
                        continue                                            # neighbor_outside never matches
                    if q[k + dk - q_lo[2], j + dj - q_lo[1], i + di - q_lo[0]] == c:
                        same = True
                        break
                if not same:
                    q[k - q_lo[2], j - q_lo[1], i - q_lo[0]] = 1 if c == 0 else 0
    return q


@pytest.mark.parametrize("shape,seed", [((9, 11, 14), 1), ((16, 8, 12), 2), ((6, 6, 6), 3)])
def test_tortuosity_remspot_matches_sequential_order(capi, shape, seed):
    """Salt-and-pepper noise (chains of mutually dependent isolated voxels) so the in-place order matters; the whole
    domain as one box against the oracle, then tile by tile against the literal loop on the same ghosted array."""
    from oracle import oi_numpy as o
    rng = np.random.default_rng(seed)
    ph = (rng.random(shape) < 0.5).astype(np.int32)
    nz, ny, nx = shape
    dom_hi = (nx - 1, ny - 1, nz - 1)
    out = capi.ref_tortuosity_remspot(ph, (0, 0, 0), (0, 0, 0), dom_hi, (0, 0, 0), dom_hi)
    assert np.array_equal(out, o.remspot(ph, 1))
    assert not np.array_equal(out, ph)
    # tiles of a one-ghost array, applied one after the other as the reference's MFIter loop does
    qg = np.zeros((nz + 2, ny + 2, nx + 2), dtype=np.int32)
    qg[1:-1, 1:-1, 1:-1] = ph
    want = qg.copy()
    for lo, hi in _tiles(shape, (4, 5, 1024000)):
        qg = capi.ref_tortuosity_remspot(qg, (-1, -1, -1), lo, hi, (0, 0, 0), dom_hi)
        want = _remspot_box_numpy(want, (-1, -1, -1), lo, hi, (0, 0, 0), dom_hi)
        assert np.array_equal(qg, want)
