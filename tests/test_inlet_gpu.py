"""GPU: the two inflow-sample search kernels (csrc/lbm_inlet.cuh) through the C ABI against numpy, in float32 with the kernels' operation order:
luw_inlet_nearest == first argmin of (dx*dx + dy*dy) + dz*dz (NearestNeighborInterpolator::eval, FX/interpolation.cpp:53-62), luw_inlet_knn == the 64 smallest
in-plane distances, their maximum, the first coincident sample (the selection loop of KNNInterpolatorHD::eval, FX/interpolation_hd.cpp:232-296). Bit-identity of the
SLOT ORDER and of the velocities built from it is the business of baseline/inlet_parity.cpp (against the reference's own code; tests/test_reference_driver.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    from latticeurbanwind_b200 import _cabi as A
    return A, A.lib()


@pytest.mark.parametrize("ncells,npts", [(5000, 3000), (1, 1), (130, 0)], ids=["5000x3000", "1x1", "no-samples"])
def test_nearest_sample_search_equals_numpy(ncells, npts):
    A, L = _lib()
    rng = np.random.default_rng(11)
    cell = rng.integers(-40, 40, (3, ncells)).astype(np.float32) + np.float32(0.5)  # lattice positions: many equal distances
    pts = (rng.integers(-20, 20, (npts, 3)) * 2).astype(np.float32) + np.float32(0.5)
    near = np.full(ncells, 7, np.uint32)
    A.check(L.luw_inlet_nearest(0, ncells, cell.ctypes.data, npts, pts.ctypes.data, near.ctypes.data))
    if npts == 0:
        assert np.all(near == 0xFFFFFFFF)
        return
    d = cell.T[:, None, :] - pts[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    assert d2.dtype == np.float32
    assert np.array_equal(near, np.argmin(d2, axis=1).astype(np.uint32))  # argmin returns the first minimum, like the reference's `d2 < best`


@pytest.mark.parametrize("npts", [4000, 64, 17, 0], ids=["4000", "64", "17", "none"])
def test_k_nearest_selection_equals_numpy(npts):
    A, L = _lib()
    rng = np.random.default_rng(13)
    ncells = 3000
    cell = rng.uniform(-100, 100, (2, ncells)).astype(np.float32)
    pts = rng.uniform(-100, 100, (npts, 2)).astype(np.float32)
    if npts >= 64:
        pts[37] = cell[:, 5]; pts[50] = cell[:, 5]  # two samples on top of cell 5: the first one is reported
    kept = np.full((ncells, 64), 9, np.uint32); used = np.zeros(ncells, np.uint32); mr = np.zeros(ncells, np.float32); ex = np.zeros(ncells, np.int32)
    A.check(L.luw_inlet_knn(0, ncells, cell.ctypes.data, npts, pts.ctypes.data, kept.ctypes.data, used.ctypes.data, mr.ctypes.data, ex.ctypes.data))
    for c in range(ncells):
        if npts == 0:
            assert used[c] == 0 and ex[c] == -1
            continue
        s1 = pts[:, 0] - cell[0, c]; s2 = pts[:, 1] - cell[1, c]
        r2 = s1 * s1 + s2 * s2
        if npts >= 64 and c == 5:
            assert ex[c] == 37
            continue
        assert ex[c] == -1 and used[c] == min(64, npts)
        got = np.sort(r2[kept[c, :used[c]]])
        assert np.array_equal(got, np.sort(r2)[:used[c]]) and mr[c] == got[-1]
        assert len(set(kept[c, :used[c]].tolist())) == used[c]
