"""Oracle restatement of getMatches.m (descriptor weighting + matchFeatures) -- TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).

getMatches.m:22-41 is restated line by line.  `matchFeatures` (Computer Vision Toolbox, closed, no version pinned by
the reference) is restated from its documentation for 'Method','Exhaustive':
  * non-binary features are normalised to unit L2 vectors before matching (the reason the reference appends a
    constant "un-normalisation" element, getMatches.m:19-27);
  * scores(i,j) = sum |f1_i - f2_j| ('SAD') or sum (f1_i - f2_j)^2 ('SSD');
  * every feature of features1 is paired with its nearest feature of features2 (first index on ties);
  * 'MatchThreshold' t (percent): pairs with score > t/100 * (largest possible score between unit vectors:
    2*sqrt(D) for SAD, 4 for SSD) are dropped;
  * 'MaxRatio' r: pairs with (nearest score)/(second nearest score) > r are dropped (if features2 has one row the
    test is skipped; a second-nearest score below 1e-6 makes the ratio 1);
  * 'Unique' true: a pair (i,j) survives only if i is also the nearest feature1 of feature2 j (forward-backward,
    first index on ties);
  * indexPairs comes out ordered by the features1 index.
The reference's drivers pass 'Method','Approximate' (completeExperiment.m:118), a randomised kd-forest whose results
cannot be reproduced; the exhaustive search is what it approximates.  PARITY UNPINNED (no MATLAB here, no fixture in
the reference).
"""
from __future__ import annotations

import numpy as np

SAD, SSD = 0, 1


def weight_descriptors(descSurface, descModel, par):
    """getMatches.m:22-41: constant extra element (UNNORMALIZE) and element-wise power (CHANGE_METRIC)."""
    dS = np.asarray(descSurface, dtype=np.float64)
    dM = np.asarray(descModel, dtype=np.float64)
    if par.get("UNNORMALIZE", False):
        # vecnorm(A, 1, 2): the 1-norm of every row (:24)
        avg_desc_len = np.mean(np.sum(np.abs(np.vstack([dS, dM])), axis=1))
        c = par["norm_factor"] * avg_desc_len
        dS = np.hstack([dS, np.full((dS.shape[0], 1), c)])          # :25
        dM = np.hstack([dM, np.full((dM.shape[0], 1), c)])          # :26
    if par.get("CHANGE_METRIC", False):
        dS = dS ** par["metric_factor"]                              # :36
        dM = dM ** par["metric_factor"]                              # :37
    return dS, dM


def normalize_rows(X):
    """matchFeatures' unit-vector normalisation: x / (||x||_2 + eps)."""
    n = np.sqrt(np.sum(X * X, axis=1, keepdims=True))
    return X / (n + np.finfo(np.float64).eps)


def score_matrix(F1, F2, metric, chunk=64):
    out = np.empty((F1.shape[0], F2.shape[0]), dtype=np.float64)
    for a in range(0, F1.shape[0], chunk):
        d = F1[a:a + chunk, None, :] - F2[None, :, :]
        out[a:a + chunk] = np.abs(d).sum(axis=2) if metric == SAD else (d * d).sum(axis=2)
    return out


def match_features_exhaustive(F1, F2, metric=SAD, match_threshold=1.0, max_ratio=0.6, unique=False, return_all=False):
    """matchFeatures(F1, F2, 'Method','Exhaustive', ...) -> (indexPairs [P,2] 0-based, matchMetric [P]).
    return_all adds a dict with the score matrix and the per-row decisions (used by the tests to find near-ties)."""
    F1 = normalize_rows(np.asarray(F1, dtype=np.float64))
    F2 = normalize_rows(np.asarray(F2, dtype=np.float64))
    n1, n2 = F1.shape[0], F2.shape[0]
    D = F1.shape[1]
    S = score_matrix(F1, F2, metric)
    j1 = np.argmin(S, axis=1)                                       # first index on ties
    d1 = S[np.arange(n1), j1]
    max_val = 2.0 * np.sqrt(D) if metric == SAD else 4.0
    thr = match_threshold * 0.01 * max_val
    keep_thr = d1 <= thr
    if n2 > 1:
        S2 = S.copy()
        S2[np.arange(n1), j1] = np.inf
        d2 = S2.min(axis=1)
        zero = d2 < 1e-6
        ratio = np.where(zero, 1.0, d1 / np.where(zero, 1.0, d2))
        keep_ratio = ratio <= max_ratio
    else:
        d2 = np.full(n1, np.inf)
        ratio = np.zeros(n1)
        keep_ratio = np.ones(n1, dtype=bool)
    back = np.argmin(S, axis=0)                                     # nearest feature1 of every feature2
    keep_unique = (back[j1] == np.arange(n1)) if unique else np.ones(n1, dtype=bool)
    keep = keep_thr & keep_ratio & keep_unique
    rows = np.nonzero(keep)[0]
    pairs = np.stack([rows, j1[rows]], axis=1).astype(np.int64)
    if return_all:
        return pairs, d1[rows], dict(S=S, j1=j1, d1=d1, d2=d2, thr=thr, ratio=ratio, back=back, keep=keep)
    return pairs, d1[rows]


def getMatches(descSurface, descModel, par, return_metric=False, return_all=False):
    """getMatches.m:1-56 -> matches [P,2] (0-based indices into descSurface / descModel)."""
    dS, dM = weight_descriptors(descSurface, descModel, par)
    metric = SSD if str(par.get("Metric", "SSD")).upper() == "SSD" else SAD
    res = match_features_exhaustive(dS, dM, metric, par.get("MatchThreshold", 1.0), par.get("MaxRatio", 0.6),
                                    bool(par.get("Unique", False)), return_all=return_all)
    if return_all:
        return res
    return res if return_metric else res[0]
