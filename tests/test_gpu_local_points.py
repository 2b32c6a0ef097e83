"""Batched getLocalPoints (getLocalPoints.m:5-36) on the GPU against the oracle restatement: membership, order
(original model order), relative coordinates and distances bit-exact (FP64, same operation order)."""
import numpy as np
import pytest

import oracle
from pcreg_b200 import synth

pytestmark = pytest.mark.gpu


def test_local_points_match_oracle(pcreg):
    model = np.asarray(synth.make_model(60_000, 3), dtype=np.float64)
    g = synth.rng(4)
    centres = np.vstack([model[g.integers(0, model.shape[0], 40)] + g.normal(0, 0.5, (40, 3)), g.uniform(-50, 150, (10, 3))])
    m = pcreg.Model(model)
    got = pcreg.getLocalPoints_batch(m, centres, 3.5, 30, 6000, return_idx=True)
    n_empty = 0
    for c, (p, d, idx) in zip(centres, got):
        wp, wd = oracle.getLocalPoints(model, 3.5, c, 30, 6000)
        if wp is None:
            assert p is None and d is None
            n_empty += 1
        else:
            assert np.array_equal(p, wp) and np.array_equal(d, wd)
            assert np.all(np.diff(idx) > 0) and np.array_equal(model[idx] - c, wp)
    assert 0 < n_empty < len(centres)
    m.destroy()


def test_local_points_debug_script_setting(pcreg):
    """getLocalPointsDebug.m:2-18: rand(1e5,3)*10, c = [3,3,3], R = 2.5 -- v1 == v2 == GPU."""
    g = synth.rng(2)
    pts = g.uniform(0, 10, (100_000, 3))
    a, da = pcreg.getLocalPoints(pts, 2.5, [3, 3, 3], 1, np.inf)
    b, db = oracle.getLocalPoints_v2(pts, 2.5, [3, 3, 3], 1, np.inf)
    assert np.array_equal(a, b) and np.array_equal(da, db) and np.all(da < 2.5)
    assert pcreg.getLocalPoints(pts, 2.5, [3, 3, 3], 10 ** 6, np.inf) == (None, None)
    assert pcreg.getLocalPoints(pts, 2.5, [3, 3, 3], 1, 10) == (None, None)


def test_local_points_boundary_is_strict(pcreg):
    """Points at distance exactly R are excluded (dists < R, getLocalPoints.m:25)."""
    pts = np.array([[1.0, 0, 0], [0, 2.0, 0], [0, 0, 2.0], [0, 0, 1.9999999999999998], [3.0, 4.0, 0.0], [0.6, 0.8, 0.0]])
    p, d = pcreg.getLocalPoints(pts, 2.0, [0, 0, 0], 0, np.inf)
    wp, wd = oracle.getLocalPoints(pts, 2.0, [0, 0, 0], 0, np.inf)
    assert np.array_equal(p, wp) and p.shape[0] == 3
    p, d = pcreg.getLocalPoints(pts, 5.0, [0, 0, 0], 0, np.inf)
    assert p.shape[0] == 5                        # the 3-4-5 point sits exactly on the sphere: excluded


def test_local_points_feed_align_points(pcreg):
    """The neighbourhoods are exactly what AlignPoints* consume in the reference pipeline
    (getSpacialHistogramDescriptors.m:50-93): align them in one batched call and compare with the oracle."""
    model = np.asarray(synth.make_model(80_000, 5), dtype=np.float64)
    g = synth.rng(6)
    centres = model[g.integers(0, model.shape[0], 12)]
    m = pcreg.Model(model)
    nbs = [p for p, d in pcreg.getLocalPoints_batch(m, centres, 3.5, 50, 6000) if p is not None]
    assert len(nbs) >= 8
    for p, (a, cu, c) in zip(nbs, pcreg.align_points_batch(pcreg.ALIGN_KNN_FRAC, nbs)):
        wa, wcu, wc = oracle.AlignPoints_KNN(p)
        assert np.max(np.abs(cu - wcu)) < 1e-9 and np.max(np.abs(a - wa)) < 1e-9 * max(1.0, np.abs(p).max())
    m.destroy()
