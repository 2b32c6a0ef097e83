"""getMatches.m on the GPU (match.cu, through pcreg_get_matches) against the oracle restatement (oracle/matching.py).

Parity rule (BASELINE.json north_star, the one used for correspondences): index decisions are compared exactly
wherever the oracle's own decision margins exceed 1e-9 relative -- the two sides sum 981 terms in different orders and
CUDA's pow is not correctly rounded, so scores agree to ~1e-15 relative, not bit for bit.  Scores: rtol 1e-12."""
import numpy as np
import pytest

import oracle
from pcreg_b200 import synth

pytestmark = pytest.mark.gpu

PAR = dict(UNNORMALIZE=True, norm_factor=2, CHANGE_METRIC=True, metric_factor=0.6, Method="Approximate",
           MatchThreshold=10, MaxRatio=0.99, Metric="SAD", Unique=True)          # completeExperiment.m:112-122
MARGIN = 1e-9


def _descriptors(n, dim, seed, base=None, noise=0.0):
    """Count-like descriptors (what histcn produces): Poisson counts around smooth random profiles."""
    g = np.random.default_rng(seed)
    if base is None:
        base = g.gamma(0.6, 4.0, (n, dim))
    lam = base * (1.0 + noise * g.standard_normal(base.shape)).clip(0.0, None)
    return g.poisson(lam).astype(np.float64), base


def _decided_rows(info, par):
    """Rows of descSurface whose every matchFeatures decision has a margin above MARGIN in the oracle."""
    S, j1, d1, d2 = info["S"], info["j1"], info["d1"], info["d2"]
    n1 = S.shape[0]
    ok = np.ones(n1, dtype=bool)
    if S.shape[1] > 1:
        ok &= np.abs(d2 - d1) > MARGIN * np.maximum(d2, 1e-300)                  # which model row is the nearest
    ok &= np.abs(d1 - info["thr"]) > MARGIN * info["thr"]                        # MatchThreshold
    if S.shape[1] > 1:
        ok &= np.abs(info["ratio"] - par.get("MaxRatio", 0.6)) > MARGIN           # MaxRatio
        ok &= np.abs(d2 - 1e-6) > 1e-12
    if par.get("Unique", False):
        col = S[:, j1]                                                           # column of every row's nearest model row
        srt = np.sort(col, axis=0)
        if n1 > 1:
            ok &= (srt[1] - srt[0]) > MARGIN * np.maximum(srt[1], 1e-300)        # who is nearest from the other side
    return ok


def _check(pcreg, dS, dM, par, min_decided=0.99, min_matches=1):
    wp, wm, info = oracle.getMatches(dS, dM, par, return_all=True)
    gp, gm = pcreg.getMatches(dS, dM, par, return_metric=True)
    assert gp.ndim == 2 and gp.shape[1] == 2 and gm.shape == (gp.shape[0],)
    assert np.all(np.diff(gp[:, 0]) > 0), "pairs must come out ordered by the surface row"
    decided = _decided_rows(info, par)
    assert decided.mean() >= min_decided, "test data too degenerate: %.3f decided" % decided.mean()
    w = {int(i): (int(j), float(m)) for (i, j), m in zip(wp, wm)}
    g = {int(i): (int(j), float(m)) for (i, j), m in zip(gp, gm)}
    for i in np.nonzero(decided)[0]:
        i = int(i)
        assert (i in w) == (i in g), "row %d: oracle %s, GPU %s" % (i, i in w, i in g)
        if i in w:
            assert w[i][0] == g[i][0], "row %d matched to %d, oracle says %d" % (i, g[i][0], w[i][0])
            assert abs(w[i][1] - g[i][1]) <= 1e-12 * max(abs(w[i][1]), 1e-3)
    assert len(w) >= min_matches, "vacuous test: the oracle found %d matches" % len(w)
    return wp, gp


def test_getmatches_reference_parameters(pcreg):
    dM, base = _descriptors(1200, 980, 3)
    dS, _ = _descriptors(300, 980, 4, base=base[::4], noise=0.05)               # surface = noisy re-draws of every 4th model row
    wp, gp = _check(pcreg, dS, dM, PAR, min_matches=100)
    assert np.mean(gp[:, 1] == 4 * gp[:, 0]) > 0.9                              # and they are the right ones


@pytest.mark.parametrize("override,dim,n1,n2", [
    (dict(Metric="SSD", MatchThreshold=30.0, MaxRatio=0.9), 980, 130, 700),
    (dict(Unique=False), 980, 257, 513),
    (dict(UNNORMALIZE=False, MatchThreshold=18.0), 980, 100, 333),
    (dict(CHANGE_METRIC=False, MatchThreshold=25.0), 980, 100, 333),
    (dict(UNNORMALIZE=False, CHANGE_METRIC=False, Metric="SSD", MatchThreshold=50.0, Unique=False), 33, 65, 64),
    (dict(MatchThreshold=100.0, MaxRatio=1.0), 15, 1, 300),                       # one surface descriptor
    (dict(MatchThreshold=100.0), 16, 70, 1),                                      # one model descriptor: no ratio test
    (dict(MatchThreshold=6.0), 980, 200, 400),                                    # the threshold rejects about half
])
def test_getmatches_variants(pcreg, override, dim, n1, n2):
    par = dict(PAR, **override)
    dM, base = _descriptors(n2, dim, 10 + n1)
    k = max(1, n2 // n1)
    sel = (np.arange(n1) * k) % n2
    dS, _ = _descriptors(n1, dim, 20 + n2, base=base[sel], noise=0.1)
    _check(pcreg, dS, dM, par)


def test_threshold_rejects_something(pcreg):
    dM, base = _descriptors(400, 980, 7)
    dS, _ = _descriptors(200, 980, 8, base=base[:200], noise=0.3)
    loose = pcreg.getMatches(dS, dM, dict(PAR, MatchThreshold=100.0, MaxRatio=1.0, Unique=False))
    tight = pcreg.getMatches(dS, dM, dict(PAR, MatchThreshold=6.0, MaxRatio=1.0, Unique=False))
    assert loose.shape[0] == 200 and 0 < tight.shape[0] < 200


def test_exact_ties_take_the_first_index(pcreg):
    """Duplicated model rows give bit-identical scores on both sides: the nearest index is the smaller one, the ratio
    d1/d2 = 1 is rejected by MaxRatio 0.99 and accepted by 1.0; duplicated surface rows: Unique keeps the first."""
    dM, _ = _descriptors(90, 980, 31)
    dM = np.vstack([dM, dM[:30]])                      # model rows 90..119 duplicate rows 0..29
    dS = np.vstack([dM[:30], dM[5:6]])                 # surface row 30 duplicates surface row 5
    par = dict(PAR, MatchThreshold=100.0, MaxRatio=1.0, Unique=False)
    gp, gm = pcreg.getMatches(dS, dM, par, return_metric=True)
    assert np.array_equal(gp[:, 0], np.arange(31))
    assert np.array_equal(gp[:, 1], np.r_[np.arange(30), 5]) and np.all(gm == 0.0)
    # identical nearest and second nearest (both 0 < 1e-6): matchFeatures sets the ratio to 1 -> rejected at 0.99
    assert pcreg.getMatches(dS, dM, dict(par, MaxRatio=0.99)).shape[0] == 0
    gu = pcreg.getMatches(dS, dM, dict(par, Unique=True))
    assert np.array_equal(gu[:, 0], np.arange(30))     # surface row 30 loses model row 5 to surface row 5
    wp = oracle.getMatches(dS, dM, dict(par, Unique=True))
    assert np.array_equal(gu, wp)


def test_empty_inputs(pcreg):
    dM, _ = _descriptors(10, 980, 1)
    assert pcreg.getMatches(np.zeros((0, 980)), dM, PAR).shape == (0, 2)
    assert pcreg.getMatches(dM, np.zeros((0, 980)), PAR).shape == (0, 2)


def test_self_match_at_full_size(pcreg):
    """Size-independent property at a driver-sized problem (3000 surface x 20000 model descriptors, 981 dimensions, more
    than one column chunk): the model set contains a shuffled copy of the surface set, so every surface descriptor has
    a zero-score partner, every pair is mutual, and the ratio test passes (second nearest > 0)."""
    n1, n2 = 3000, 20000
    dS, _ = _descriptors(n1, 980, 41)
    extra, _ = _descriptors(n2 - n1, 980, 42)
    g = np.random.default_rng(43)
    perm = g.permutation(n2)
    dM = np.vstack([dS, extra])[perm]
    where = np.empty(n2, dtype=np.int64)
    where[perm] = np.arange(n2)
    pcreg.set_profiling(True)
    gp, gm = pcreg.getMatches(dS, dM, PAR, return_metric=True)
    prof = pcreg.last_profile()
    pcreg.set_profiling(False)
    assert np.array_equal(gp[:, 0], np.arange(n1))
    assert np.array_equal(gp[:, 1], where[:n1])
    assert np.all(gm == 0.0)
    assert prof["match_terms"] == float(n1) * n2 * 981 and prof["match_score_ms"] > 0.0
    print("k_match_scores: %.2f ms, %.2f T terms/s" % (prof["match_score_ms"], prof["match_terms"] / prof["match_score_ms"] / 1e9))


def test_matches_from_real_descriptors(pcreg):
    """Keypoints -> spherical histograms -> getMatches, everything on the GPU, against the oracle given the same
    descriptors; a noisy copy of the model's keypoints must mostly match back to itself."""
    model = np.asarray(synth.make_model(150_000, 17), dtype=np.float64)
    g = synth.rng(18)
    kpM = model[g.choice(model.shape[0], 160, replace=False)]
    kpS = kpM[:60] + g.normal(0, 0.02, (60, 3))
    opts = dict(min_pts=100, max_pts=np.inf, R=3.5, thVar=(1.0, 1.0), k=0.85, ALIGN_POINTS=True)
    m = pcreg.Model(model)
    fM, dM = pcreg.getSpacialHistogramDescriptors(m, kpM, opts)
    fS, dS = pcreg.getSpacialHistogramDescriptors(m, kpS, opts)
    m.destroy()
    assert dM.shape[0] > 100 and dS.shape[0] > 40
    wp, gp = _check(pcreg, dS, dM, PAR, min_decided=0.9, min_matches=10)
    good = np.linalg.norm(fS[gp[:, 0]] - fM[gp[:, 1]], axis=1) < 0.2
    assert good.mean() > 0.8


def test_transfer_colors(pcreg):
    """ColorCodeModel.m:12-18 as one batched 1-NN search."""
    g = np.random.default_rng(5)
    col_pts = g.uniform(-10, 10, (5000, 3))
    colors = g.integers(0, 256, (5000, 3)).astype(np.uint8)
    pts = g.uniform(-10, 10, (2000, 3))
    m = pcreg.Model(col_pts)
    got = pcreg.transfer_colors(m, pts, colors)
    m.destroy()
    want = colors[oracle.nn_brute(col_pts, pts)[0]]
    assert np.array_equal(got, want)
