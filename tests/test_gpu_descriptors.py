"""getSpacialHistogramDescriptors.m on the GPU (descriptor.cu) against the oracle restatement (oracle/descriptors.py):
the set of surviving keypoints and their 980-bin histograms.  Counts are integers: compared exactly."""
import numpy as np
import pytest

import oracle
from pcreg_b200 import synth

pytestmark = pytest.mark.gpu


def _setup(nm, nk, seed):
    model = np.asarray(synth.make_model(nm, seed), dtype=np.float64)
    g = synth.rng(seed + 1)
    on = model[g.integers(0, nm, nk - nk // 4)] + g.normal(0, 0.3, (nk - nk // 4, 3))     # near the surface
    off = g.uniform(model.min(0), model.max(0), (nk // 4, 3))                              # mostly empty space -> []
    return model, np.vstack([on, off])


@pytest.mark.parametrize("k,align,thvar", [("all", True, (1.0, 1.0)), (0.85, True, (1.0, 1.0)), (0.85, True, (1.2, 1.5)),
                                           ("all", False, (1.0, 1.0)), (0.5, False, (1.3, 1.0))])
def test_spatial_histogram_matches_oracle(pcreg, k, align, thvar):
    model, kp = _setup(120_000, 40, 11)
    opts = dict(min_pts=150, max_pts=2500, R=3.5, thVar=thvar, k=k, ALIGN_POINTS=align)
    wf, wd = oracle.getSpacialHistogramDescriptors(model, kp, opts)
    m = pcreg.Model(model)
    gf, gd, status, counts = pcreg.getSpacialHistogramDescriptors(m, kp, opts, return_status=True)
    m.destroy()
    assert wf.shape[0] > 5 and np.sum(status == 1) > 0
    if thvar != (1.0, 1.0):
        assert np.sum(status == 2) > 0          # the variance rejection really rejects something
    assert np.array_equal(gf, wf), "different set of surviving keypoints"
    assert gd.shape == wd.shape == (wf.shape[0], 980)
    assert np.array_equal(gd, wd), "%d histogram bins differ" % int(np.sum(gd != wd))
    # every point of a neighbourhood lands in exactly one bin unless it coincides with the keypoint (theta = NaN)
    assert np.all(gd.sum(axis=1) <= counts[status == 0]) and np.all(gd.sum(axis=1) >= counts[status == 0] - 1)


def test_phi_quirk_and_histcounts_edges(pcreg):
    """phi = atan2(y, y) (getSpacialHistogramDescriptors.m:152) only reaches the phi bins of pi/4 and -3pi/4 (and 0 for
    y == 0); a point exactly on the outer radius edge is excluded by getLocalPoints' strict `< R` before histcn sees it."""
    model, kp = _setup(60_000, 12, 5)
    opts = dict(min_pts=50, max_pts=np.inf, R=3.0, thVar=(1.0, 1.0), k="all", ALIGN_POINTS=True)
    m = pcreg.Model(model)
    gf, gd = pcreg.getSpacialHistogramDescriptors(m, kp, opts)
    m.destroy()
    h = gd.reshape(-1, 14, 7, 10)              # reshape(counts, [], 1) of (r, theta, phi): phi is the slowest index
    used = np.nonzero(h.sum(axis=(0, 2, 3)))[0]
    assert set(used.tolist()) <= {1, 7, 8}     # -3pi/4 -> bin 2, 0 -> bin 8, pi/4 -> bin 9 (1-based)
    assert {1, 8} <= set(used.tolist())


def test_keypoint_on_a_model_point_drops_it(pcreg):
    """A model point that coincides with the keypoint has r = 0, theta = acos(0/0) = NaN: histcn drops it (histcn.m:125)."""
    model, _ = _setup(50_000, 4, 9)
    kp = model[[10, 2000, 30000]]
    opts = dict(min_pts=50, max_pts=np.inf, R=3.5, thVar=(1.0, 1.0), k="all", ALIGN_POINTS=False)
    m = pcreg.Model(model)
    gf, gd, status, counts = pcreg.getSpacialHistogramDescriptors(m, kp, opts, return_status=True)
    m.destroy()
    wf, wd = oracle.getSpacialHistogramDescriptors(model, kp, opts)
    assert np.array_equal(gd, wd)
    assert np.all(gd.sum(axis=1) == counts[status == 0] - 1)


def test_spatial_histogram_on_a_model_with_grid_equals_grid_less(pcreg):
    """The descriptor stage on a model WITH a uniform grid takes its neighbourhoods from the grid walk of local_points.cu
    (device-resident variant): same surviving keypoints, same histograms as on the grid-less model and as the oracle."""
    model, kp = _setup(300_000, 1500, 21)
    opts = dict(min_pts=150, max_pts=6000, R=3.5, thVar=(1.1, 1.2), k=0.85, ALIGN_POINTS=True)
    mg = pcreg.Model(model, grid=True, voxel_map=-1)
    mb = pcreg.Model(model)
    a = pcreg.getSpacialHistogramDescriptors(mg, kp, opts, return_status=True)
    b = pcreg.getSpacialHistogramDescriptors(mb, kp, opts, return_status=True)
    mg.destroy(); mb.destroy()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert a[0].shape[0] > 300
    sel = np.arange(0, kp.shape[0], 97)
    wf, wd = oracle.getSpacialHistogramDescriptors(model, kp[sel], opts)
    keep = np.nonzero(a[2][sel] == 0)[0]
    assert np.array_equal(wf, kp[sel][keep])
    rows = np.cumsum(a[2] == 0) - 1
    assert np.array_equal(wd, a[1][rows[sel][keep]])
