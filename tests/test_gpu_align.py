"""Parity of the batched AlignPoints* kernel against the oracle restatement of the six .m functions.
Tolerance: coeff 1e-9 absolute (same algorithm in FP64, different summation order / eigen-solver),
aligned points 1e-9 relative to the cloud extent; checkAlignment (visualizeGTMatches.m:417-421) < 1e-6."""
import numpy as np
import pytest

import oracle
from pcreg_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _cmp(got, want, scale):
    ga, gc = got[0], got[1]
    wa, wc = want[0], want[1]
    if wa is None:
        assert ga is None and gc is None
        return
    assert gc is not None
    assert np.max(np.abs(gc - wc)) < TOL, "coeff differs by %g" % np.max(np.abs(gc - wc))
    assert oracle.check_alignment(gc, wc) < 1e-6
    assert np.max(np.abs(ga - wa)) < TOL * scale
    if len(want) > 2:
        assert np.max(np.abs(got[2] - want[2])) < 1e-12 * scale


CASES = [
    ("AlignPoints", lambda o, p: o.AlignPoints(p), lambda g, p: g.AlignPoints(p)),
    ("AlignPoints_KNN", lambda o, p: o.AlignPoints_KNN(p), lambda g, p: g.AlignPoints_KNN(p)),
    ("AlignPoints_KNN_C1", lambda o, p: o.AlignPoints_KNN(p, True, False), lambda g, p: g.AlignPoints_KNN(p, True, False)),
    ("AlignPoints_KNN_C2", lambda o, p: o.AlignPoints_KNN(p, False, True), lambda g, p: g.AlignPoints_KNN(p, False, True)),
    ("AlignPoints_KNN_C1C2", lambda o, p: o.AlignPoints_KNN(p, True, True), lambda g, p: g.AlignPoints_KNN(p, True, True)),
    ("AlignPoints_knn_500", lambda o, p: o.AlignPoints_knn(p, 500), lambda g, p: g.AlignPoints_knn(p, 500)),
    ("AlignPoints_knn_1500", lambda o, p: o.AlignPoints_knn(p, 1500), lambda g, p: g.AlignPoints_knn(p, 1500)),
    ("AlignPoints_weighted", lambda o, p: o.AlignPoints_weighted(p), lambda g, p: g.AlignPoints_weighted(p)),
    ("AlignPoints_c", lambda o, p: o.AlignPoints_c(p)[:2], lambda g, p: g.AlignPoints_c(p)),
    ("AlignPoints_KNN_c", lambda o, p: o.AlignPoints_KNN_c(p), lambda g, p: g.AlignPoints_KNN_c(p)),
]


@pytest.mark.parametrize("name,ofn,gfn", CASES, ids=[c[0] for c in CASES])
def test_align_variants_double(pcreg, name, ofn, gfn):
    for p in synth.make_neighbourhoods(5, 42):
        want = ofn(oracle, p)
        got = gfn(pcreg, p)
        _cmp(got, want, np.abs(p).max())


def test_align_single_class_roundtrip(pcreg):
    """Class single in -> class single out (reference clouds are single: upsampleMesh.m:21); values
    equal the FP64 oracle on the same (float-valued) inputs to float rounding."""
    for p in synth.make_neighbourhoods(3, 43, dtype=np.float32):
        a, cu, c = pcreg.AlignPoints_KNN(p)
        wa, wcu, wc = oracle.AlignPoints_KNN(p.astype(np.float64))
        assert a.dtype == np.float32
        assert np.max(np.abs(cu - wcu)) < TOL
        assert np.max(np.abs(a.astype(np.float64) - wa)) <= 1e-6 * np.abs(p).max()


def test_align_returns_empty_like_reference(pcreg):
    """AlignPoints_c / _KNN_c return [] when fewer than 25 points fall inside r = 2.0 (AlignPoints_c.m:16-18)."""
    g = synth.rng(3)
    d = g.standard_normal((400, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p = d * g.uniform(2.5, 3.4, (400, 1)) + np.array([10.0, 20.0, 30.0])       # a hollow shell: nothing near the centroid
    a, cu = pcreg.AlignPoints_c(p)
    assert a is None and cu is None
    assert oracle.AlignPoints_c(p)[0] is None
    a, cu, c = pcreg.AlignPoints_KNN_c(p)
    assert a is None and cu is None
    np.testing.assert_allclose(c, p.mean(axis=0), rtol=0, atol=1e-12)


def test_align_batch_ragged(pcreg):
    """Batched call, ragged neighbourhood sizes incl. tiny ones; identical to one-by-one calls."""
    nbs = synth.make_neighbourhoods(12, 44, nmin=30, nmax=3000)
    batch = pcreg.align_points_batch(pcreg.ALIGN_KNN_FRAC, nbs)
    for p, (a, cu, c) in zip(nbs, batch):
        wa, wcu, wc = oracle.AlignPoints_KNN(p)
        _cmp((a, cu, c), (wa, wcu, wc), np.abs(p).max())


def test_align_duplicate_distances_stable_selection(pcreg):
    """Mirror-symmetric cloud: many exactly equal centroid distances at the 85 % boundary."""
    g = synth.rng(8)
    half = g.normal(0, 1.0, (300, 3)) * np.array([3.0, 1.5, 0.4])
    p = np.vstack([half, -half]) + np.array([5.0, -7.0, 2.0])
    p = np.vstack([p, p[:101]])                     # exact duplicates as well
    got = pcreg.AlignPoints_KNN(p)
    want = oracle.AlignPoints_KNN(p)
    _cmp(got, want, np.abs(p).max())


def test_align_rotation_equivariance_property(pcreg):
    """Size-independent property: aligning a rotated copy gives the same aligned cloud (up to the
    vote-defined signs), i.e. the local reference frame is rotation invariant."""
    p = synth.make_neighbourhoods(1, 45)[0]
    R = synth.rot_xyz([0.7, -1.1, 2.0])
    a1, c1 = pcreg.AlignPoints(p)
    a2, c2 = pcreg.AlignPoints(p @ R)
    assert np.max(np.abs(np.abs(a1) - np.abs(a2))) < 1e-8 * np.abs(p).max()
