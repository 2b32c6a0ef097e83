"""Oracle restatement of the six AlignPoints* PCA local-reference-frame functions.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  MATLAB's pca() is restated from its
documentation / shipped source behaviour:
  pca(X,'Algorithm','eig'[, 'Centered','off']) = eigen-decomposition of Xc'*Xc/(n-1)
  (/n uncentred), eigenvalues DESCENDING, score = Xc*coeff, and the sign convention
  "the largest-magnitude element of every coeff column is positive" (first max on ties).
MATLAB is not installed here, so these are parity-unpinned beyond algebraic identities.
"""
from __future__ import annotations

import numpy as np

from .primitives import getLocalPoints, matlab_round


def _vecnorm_rows(p):
    return np.sqrt((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]) + p[:, 2] * p[:, 2])


def pca_eig(X, centered: bool = True):
    """MATLAB pca(X,'Algorithm','eig'[,'Centered','off']) -> (coeff, score, mu)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    if centered:
        mu = X.mean(axis=0)
        Xc = X - mu
        dof = max(n - 1, 1)
    else:
        mu = np.zeros(3)
        Xc = X
        dof = max(n, 1)
    C = (Xc.T @ Xc) / dof
    C = 0.5 * (C + C.T)
    w, V = np.linalg.eigh(C)            # ascending
    order = np.argsort(-w, kind="stable")
    coeff = V[:, order]
    # sign convention: largest |element| of each column positive (pca.m)
    for c in range(3):
        i = int(np.argmax(np.abs(coeff[:, c])))
        if coeff[i, c] < 0:
            coeff[:, c] = -coeff[:, c]
    score = Xc @ coeff
    return coeff, score, mu


def _disambiguate(coeff, pts_lrf, k):
    """AlignPoints.m:10-25 (identical in every variant): majority vote on the sign of score
    columns 1 and 3 against threshold k/2, y sign = det (a float ~ +-1, NOT snapped)."""
    x_sign = float(np.sum(pts_lrf[:, 0] > 0) >= k / 2.0) * 2 - 1          # :14,18
    z_sign = float(np.sum(pts_lrf[:, 2] > 0) >= k / 2.0) * 2 - 1          # :15,19
    y_sign = np.linalg.det(coeff * np.array([x_sign, 1.0, z_sign]))       # :22
    return coeff * np.array([x_sign, y_sign, z_sign])                     # :25


def AlignPoints(pts):
    """AlignPoints.m:1-29 -> (pts_aligned, coeff_unambig)."""
    pts = np.asarray(pts, dtype=np.float64)
    coeff, lrf, _ = pca_eig(pts)                                          # :6
    cu = _disambiguate(coeff, lrf, pts.shape[0])
    return pts @ cu, cu                                                   # :28 (uncentred pts)


def _k_nearest_to_centroid(pts, K):
    c = pts.mean(axis=0)                                                  # AlignPoints_KNN.m:17
    rel = pts - c                                                         # :22
    d = _vecnorm_rows(rel)                                                # :23
    I = np.argsort(d, kind="stable")                                      # :24 (MATLAB sort is stable)
    return c, rel[I[:K]]                                                  # :25-26


def AlignPoints_KNN(pts, C1: bool = False, C2: bool = False, k_frac: float = 0.85):
    """AlignPoints_KNN.m:1-60 -> (pts_aligned, coeff_unambig, c).  Reproduces the quirk that the
    vote threshold is size(pts,1)/2 (:37) while only the K selected scores are counted (:45-46)."""
    pts = np.asarray(pts, dtype=np.float64)
    N = pts.shape[0]
    K = int(matlab_round(N * k_frac))                                     # :20-21
    c, pts_k = _k_nearest_to_centroid(pts, K)
    coeff, lrf, _ = pca_eig(pts_k, centered=not C1)                       # :30-34
    if C2:
        lrf = pts @ coeff                                                 # :39-41
    cu = _disambiguate(coeff, lrf, N)                                     # :37 k = size(pts,1)
    return pts @ cu, cu, c                                                # :59


def AlignPoints_knn(pts, K):
    """AlignPoints_knn.m:1-43 -- absolute K = min(K, N) (:12)."""
    pts = np.asarray(pts, dtype=np.float64)
    N = pts.shape[0]
    K = int(min(K, N))
    c, pts_k = _k_nearest_to_centroid(pts, K)
    coeff, lrf, _ = pca_eig(pts_k)                                        # :21
    cu = _disambiguate(coeff, lrf, N)                                     # :24
    return pts @ cu, cu, c


def AlignPoints_weighted(pts, R: float = 3.5):
    """AlignPoints_weighted.m:1-49.  M = (w.*P)'*P with w = max(R - ||P||, 0) (:16-21); eig(M)
    column order is taken as ASCENDING eigenvalue (symmetric-path behaviour); because M is not
    bit-symmetric in MATLAB the reference's own order is unspecified -> parity unpinned
    (SURVEY.md section 8 a4).  No sign convention is applied to eig() output; the vote over all
    N points with threshold N/2 (:28-35) makes the result sign-invariant except on exact ties."""
    pts = np.asarray(pts, dtype=np.float64)
    c = pts.mean(axis=0)                                                  # :9
    rel = pts - c                                                         # :12
    d = _vecnorm_rows(rel)                                                # :13
    w = np.maximum(R - d, 0.0)                                            # :16-18
    M = (w[:, None] * rel).T @ rel                                        # :21
    M = 0.5 * (M + M.T)
    _, coeff = np.linalg.eigh(M)                                          # :24 ascending
    lrf = rel @ coeff                                                     # :28
    cu = _disambiguate(coeff, lrf, pts.shape[0])
    return pts @ cu, cu, c                                                # :48


def AlignPoints_c(pts, r: float = 2.0, min_local: int = 25):
    """AlignPoints_c.m:1-44 -- PCA on the points within r of the centroid; (None, None, c) where the
    reference returns [] (:16-18).  Vote threshold size(pts,1)/2 (:24)."""
    pts = np.asarray(pts, dtype=np.float64)
    c = pts.mean(axis=0)                                                  # :9
    rel, _ = getLocalPoints(pts, r, c, min_local, np.inf)                 # :13-14
    if rel is None:
        return None, None, c
    coeff, lrf, _ = pca_eig(rel)                                          # :21
    cu = _disambiguate(coeff, lrf, pts.shape[0])                          # :24
    return pts @ cu, cu, c


def AlignPoints_KNN_c(pts, k_frac: float = 0.85, r: float = 2.0, min_local: int = 25):
    """AlignPoints_KNN_c.m:1-57 -- 85 % nearest to the centroid (as relative coordinates, :14-18),
    their mean (:22), points of that subset within r of it (:26-27), PCA on those; vote threshold is
    size(pts_lrf,1)/2 here (:34), unlike AlignPoints_c."""
    pts = np.asarray(pts, dtype=np.float64)
    N = pts.shape[0]
    K = int(matlab_round(N * k_frac))                                     # :12-13
    c, pts_k = _k_nearest_to_centroid(pts, K)
    centroid = pts_k.mean(axis=0)                                         # :22
    rel, _ = getLocalPoints(pts_k, r, centroid, min_local, np.inf)        # :26-27
    if rel is None:
        return None, None, c
    coeff, lrf, _ = pca_eig(rel)                                          # :31
    cu = _disambiguate(coeff, lrf, lrf.shape[0])                          # :34
    return pts @ cu, cu, c                                                # :52
