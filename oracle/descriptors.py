"""Oracle restatement of getSpacialHistogramDescriptors.m (the 10 x 7 x 14 spherical histogram
descriptor of every keypoint's radius-R neighbourhood) and of the part of histcn.m it uses.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  MATLAB built-ins restated from documentation:
  histcounts(x, edges): bin k holds edges(k) <= x < edges(k+1), the LAST bin also holds x == edges(end),
                        anything else (outside, NaN) gets bin 0 and histcn.m:125 drops the point;
  nthroot(x, 3) = real cube root; a:d:b colon ranges are restated as a + k*d (MATLAB's colon operator may
  differ in the last ulp of an edge -- only a value exactly on an edge could tell: parity unpinned there).
The reference's `phi_spheric = atan2(y, y)` (getSpacialHistogramDescriptors.m:152) is reproduced as written:
phi takes the values pi/4 (y > 0), -3pi/4 (y < 0), 0 (y == +0) and -pi (y == -0) only.
"""
from __future__ import annotations

import numpy as np

from .align import pca_eig, _vecnorm_rows
from .primitives import getLocalPoints, matlab_round

NUM_R, NUM_THETA, NUM_PHI = 10, 7, 14          # getSpacialHistogramDescriptors.m:38-40


def histogram_edges(R, num_r=NUM_R, num_theta=NUM_THETA, num_phi=NUM_PHI):
    """getSpacialHistogramDescriptors.m:155-158."""
    r_equi = np.arange(num_r + 1) * (R ** 3 / num_r)        # 0:R^3/NUM_R:R^3
    r_bins = np.cbrt(r_equi)                                # nthroot(r_equi, 3)
    phi_bins = -np.pi + np.arange(num_phi + 1) * (2 * np.pi / num_phi)
    theta_bins = np.arange(num_theta + 1) * (np.pi / num_theta)
    return r_bins, theta_bins, phi_bins


def histcounts_bin(x, edges):
    """Third output of histcounts(x, edges): 1-based bin, 0 = not counted."""
    x = np.asarray(x, dtype=np.float64)
    edges = np.asarray(edges, dtype=np.float64)
    b = np.searchsorted(edges, x, side="right")             # edges[b-1] <= x < edges[b]
    b = np.where(x == edges[-1], len(edges) - 1, b)         # right border belongs to the last bin
    b = np.where((x < edges[0]) | (x > edges[-1]) | np.isnan(x), 0, b)
    return b.astype(np.int64)


def histcn3(X, e0, e1, e2):
    """histcn.m:97-131 for three columns with explicit edges -> counts [n0, n1, n2]."""
    loc = np.stack([histcounts_bin(X[:, 0], e0), histcounts_bin(X[:, 1], e1), histcounts_bin(X[:, 2], e2)], axis=1)
    sz = (len(e0) - 1, len(e1) - 1, len(e2) - 1)
    has = np.all(loc > 0, axis=1)                           # :125
    counts = np.zeros(sz, dtype=np.float64)
    np.add.at(counts, (loc[has, 0] - 1, loc[has, 1] - 1, loc[has, 2] - 1), 1.0)
    return counts


def spatial_histogram_of(pts_local, R, thVar=(1.0, 1.0), K="all", ALIGN_POINTS=True, edges=None):
    """The body of the second parfor (getSpacialHistogramDescriptors.m:69-175) for ONE neighbourhood that
    getLocalPoints already returned (points relative to the keypoint).  Returns the 980 counts or None
    where the reference `continue`s (variance rejection)."""
    pts_local = np.asarray(pts_local, dtype=np.float64)
    num_points = pts_local.shape[0]
    if isinstance(K, str) or K == 1:                                            # :74-75
        k = num_points
    else:
        k = int(matlab_round(num_points * K))                                   # :77
        centroid = pts_local.mean(axis=0)                                       # :79
        dists = _vecnorm_rows(pts_local - centroid)                             # :80
        I = np.argsort(dists, kind="stable")                                    # :81
        pts_local = pts_local[I]                                                # :82
    pts_k = pts_local[:k]                                                       # :84
    thVar = np.asarray(thVar, dtype=np.float64)
    coeff = pts_lrf = None
    if not (np.sum(thVar == 1) == 2) or ALIGN_POINTS:                           # :85
        coeff, pts_lrf, _ = pca_eig(pts_k)                                      # :90 (LOCAL_PCA = false)
        Xc = pts_k - pts_k.mean(axis=0)
        variances = np.sort(np.linalg.eigvalsh(0.5 * ((Xc.T @ Xc) + (Xc.T @ Xc).T) / max(k - 1, 1)))[::-1]
        with np.errstate(invalid="ignore", divide="ignore"):
            if (variances[0] / variances[1] < thVar[0]) or (variances[1] / variances[2] < thVar[1]):   # :117-120
                return None
    if ALIGN_POINTS:                                                            # :128-144
        kk = pts_lrf.shape[0]
        x_sign = float(np.sum(np.sign(pts_lrf[:, 0]) == 1) >= kk / 2.0) * 2 - 1
        z_sign = float(np.sum(np.sign(pts_lrf[:, 2]) == 1) >= kk / 2.0) * 2 - 1
        y_sign = np.linalg.det(coeff * np.array([x_sign, 1.0, z_sign]))
        coeff_unambig = coeff * np.array([x_sign, y_sign, z_sign])
        pts_local = pts_local @ coeff_unambig
    r_sph = _vecnorm_rows(pts_local)                                            # :150
    with np.errstate(invalid="ignore", divide="ignore"):
        theta = np.arccos(pts_local[:, 2] / r_sph)                              # :151
    phi = np.arctan2(pts_local[:, 1], pts_local[:, 1])                          # :152 (sic)
    if edges is None:
        edges = histogram_edges(R)
    counts = histcn3(np.stack([r_sph, theta, phi], axis=1), *edges)             # :161
    return counts.reshape(-1, order="F")                                        # :164


def getSpacialHistogramDescriptors(pts, sample_pts, options):
    """getSpacialHistogramDescriptors.m:2-183 -> (feat, desc).  options: dict with min_pts, max_pts, R, thVar,
    k ('all' or a fraction), ALIGN_POINTS."""
    pts = np.asarray(pts, dtype=np.float64)
    sample_pts = np.asarray(sample_pts, dtype=np.float64).reshape(-1, 3)
    R = float(options["R"])
    edges = histogram_edges(R)
    feat, desc = [], []
    for c in sample_pts:
        pts_local, _ = getLocalPoints(pts, R, c, options["min_pts"], options["max_pts"])     # :50,68
        if pts_local is None:
            continue
        d = spatial_histogram_of(pts_local, R, options.get("thVar", (1.0, 1.0)), options.get("k", "all"),
                                 bool(options.get("ALIGN_POINTS", True)), edges)
        if d is None:
            continue
        feat.append(c)
        desc.append(d)
    if not feat:
        return np.zeros((0, 3)), np.zeros((0, NUM_R * NUM_THETA * NUM_PHI))
    return np.stack(feat), np.stack(desc)
