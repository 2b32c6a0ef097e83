"""Oracle restatement of the reference's geometry primitives (FP64 numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line it follows.  Conventions: points are
N x 3 row-per-point, transforms are 4 x 4 ROW-VECTOR form T = [R 0; t 1] with
p' = [p 1] * T  (quickTF.m:5-7).
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------
# MATLAB built-ins restated from documentation
# ---------------------------------------------------------------------------------------------
def matlab_round(x):
    """MATLAB round(): half away from zero (numpy's round is half-to-even)."""
    x = np.asarray(x, dtype=np.float64)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def matlab_rank(A) -> int:
    """MATLAB rank(A): number of singular values > max(size(A)) * eps(norm(A))."""
    A = np.asarray(A, dtype=np.float64)
    if A.size == 0:
        return 0
    s = np.linalg.svd(A, compute_uv=False)
    tol = max(A.shape) * np.spacing(s.max())
    return int(np.sum(s > tol))


def eul2rotm(e, seq: str = "ZYX") -> np.ndarray:
    """Robotics System Toolbox eul2rotm.  Default 'ZYX': R = Rz(e1) Ry(e2) Rx(e3)
    (used at testTransformEstimation.m:12, testRANSAC.m:17); 'XYZ': R = Rx(e1) Ry(e2) Rz(e3)
    (used at pcRigidBodyTF.m:13, debugRANSAC.m:11)."""
    e = np.asarray(e, dtype=np.float64).reshape(3)

    def rx(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)

    def ry(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)

    def rz(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)

    m = {"X": rx, "Y": ry, "Z": rz}
    seq = seq.upper()
    return m[seq[0]](e[0]) @ m[seq[1]](e[1]) @ m[seq[2]](e[2])


# ---------------------------------------------------------------------------------------------
# Pose helpers
# ---------------------------------------------------------------------------------------------
def quickTF(pts, TF) -> np.ndarray:
    """quickTF.m:1-8 -- [pts 1] * TF, first three columns.

    Evaluated column by column in the fixed order ((x*T1c + y*T2c) + z*T3c) + T4c so that
    the CUDA path (same order, no FMA contraction) reproduces q bit for bit."""
    pts = np.asarray(pts, dtype=np.float64)
    TF = np.asarray(TF, dtype=np.float64)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    out = np.empty((pts.shape[0], 3), dtype=np.float64)
    for c in range(3):
        out[:, c] = ((x * TF[0, c] + y * TF[1, c]) + z * TF[2, c]) + TF[3, c]
    return out


def invertTF(TF) -> np.ndarray:
    """invertTF.m:1-8 -- [R' 0; -t R' 1]."""
    TF = np.asarray(TF, dtype=np.float64)
    Ti = np.eye(4)
    Ti[0:3, 0:3] = TF[0:3, 0:3].T
    Ti[3, 0:3] = -TF[3, 0:3] @ TF[0:3, 0:3].T
    return Ti


def pcRigidBodyTF(pts, r, t):
    """pcRigidBodyTF.m:13-19 -- T = [eul2rotm(r,'XYZ') 0; t 1], applied as p*R + t.
    Returns (pts_out, T) (the reference returns a pointCloud object; we return the array)."""
    r = np.asarray(r, dtype=np.float64).reshape(-1)
    if r.shape[0] != 3:
        raise ValueError("Rotation must be either 1x3 or 3x1 matrix.")  # pcRigidBodyTF.m:10
    T = np.eye(4)
    T[0:3, 0:3] = eul2rotm(r, "XYZ")
    T[3, 0:3] = np.asarray(t, dtype=np.float64).reshape(3)
    return quickTF(pts, T), T


# ---------------------------------------------------------------------------------------------
# Radius neighbourhood query
# ---------------------------------------------------------------------------------------------
def _vecnorm_rows(p):
    # vecnorm(x,2,2): sqrt(sum of squares) along the row, summed in column order
    return np.sqrt((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]) + p[:, 2] * p[:, 2])


def getLocalPoints(pts, R, c, min_points, max_points):
    """getLocalPoints.m:5-36 -- strict cube pre-filter, early [] if the cube holds fewer than
    min_points, strict sphere filter, points RELATIVE to c in original order, [] if the count
    is outside [min_points, max_points].  Returns (pts_sphere, dists) or (None, None) for []."""
    pts = np.asarray(pts, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64).reshape(3)
    mask = ((pts[:, 0] > c[0] - R) & (pts[:, 0] < c[0] + R)
            & (pts[:, 1] > c[1] - R) & (pts[:, 1] < c[1] + R)
            & (pts[:, 2] > c[2] - R) & (pts[:, 2] < c[2] + R))          # :8-13
    pts_cube = pts[mask]
    if pts_cube.shape[0] < min_points:                                    # :17-19
        return None, None
    pts_rel = pts_cube - c                                                # :23
    dists = _vecnorm_rows(pts_rel)                                        # :24
    m2 = dists < R                                                        # :25
    dists = dists[m2]
    pts_sphere = pts_rel[m2]
    if pts_sphere.shape[0] < min_points or pts_sphere.shape[0] > max_points:   # :31-34
        return None, None
    return pts_sphere, dists


def getLocalPoints_v2(pts, R, c, min_points, max_points):
    """getLocalPoints_v2.m:5-22 -- same sphere set without the cube pre-filter."""
    pts = np.asarray(pts, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64).reshape(3)
    pts_c = pts - c
    dists = _vecnorm_rows(pts_c)
    mask = dists < R
    n = int(mask.sum())
    if n < min_points or n > max_points:
        return None, None
    return pts_c[mask], dists[mask]


# ---------------------------------------------------------------------------------------------
# Kabsch rigid fit
# ---------------------------------------------------------------------------------------------
def estimateTransform(pts1, pts2, reflection_fix: bool = False, rank_guard: bool = True):
    """estimateTransform.m:2-74.  Returns the 4x4 row-vector T with [pts2,1]*T = [pts1,1]
    (the true direction, as the in-code comment :65 and every calcDists use say), or None
    where the reference returns [].

    reflection_fix=False reproduces the reference (R = V*U', :62, no determinant check)."""
    pts1 = np.asarray(pts1, dtype=np.float64)
    pts2 = np.asarray(pts2, dtype=np.float64)
    n = pts1.shape[0]
    if rank_guard and (matlab_rank(pts1) < 3 or matlab_rank(pts2) < 2):    # :11-14
        return None
    if n == 3:                                                            # :18-37
        c1 = pts1.mean(axis=0)
        c2 = pts2.mean(axis=0)
        n1 = np.cross(pts1[2] - pts1[1], pts1[2] - pts1[0])               # :24
        n2 = np.cross(pts2[2] - pts2[1], pts2[2] - pts2[0])               # :25
        l1 = np.median(_vecnorm_rows(pts1 - np.roll(pts1, 1, axis=0)))    # :28
        l2 = np.median(_vecnorm_rows(pts2 - np.roll(pts2, 1, axis=0)))    # :29
        p1 = c1 + (n1 / np.linalg.norm(n1)) * l1                          # :32
        p2 = c2 + (n2 / np.linalg.norm(n2)) * l2                          # :33
        pts1 = np.vstack([pts1, p1])
        pts2 = np.vstack([pts2, p2])
    d = pts1.T                                                            # :41
    m = pts2.T                                                            # :42
    cd = d.mean(axis=1, keepdims=True)                                    # :46
    cm = m.mean(axis=1, keepdims=True)                                    # :47
    H = (m - cm) @ (d - cd).T                                             # :58
    U, _, Vt = np.linalg.svd(H)                                           # :60
    V = Vt.T
    if reflection_fix and np.linalg.det(V @ U.T) < 0:
        V = V.copy()
        V[:, 2] = -V[:, 2]
    R = V @ U.T                                                           # :62
    t = cd - R @ cm                                                       # :63
    TF = np.eye(4)
    TF[0:3, 0:3] = R
    TF[0:3, 3] = t[:, 0]
    return TF.T                                                           # :71


def calcDists(T, pts1, pts2) -> np.ndarray:
    """calcDists (getInliersRANSAC.m:46-54 and six identical copies) -- SQUARED distance between
    pts1 and [pts2,1]*T."""
    pts1 = np.asarray(pts1, dtype=np.float64)
    q = quickTF(pts2, T)
    d = pts1 - q
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def ransac(pts1, pts2, coef: dict, triplets, reflection_fix: bool = False):
    """ransac.m:21-116 with the sample triplets SUPPLIED (0-based, shape [iterNum, minPtNum])
    instead of randperm(ptNum)(1:3) (:42-43) -- MATLAB's global RNG stream cannot be matched.

    coef keys as the reference's struct: thDist (compared against SQUARED distances, :49),
    thInlrRatio, REFINE.  Returns dict(T, inlierIdx (0-based), numSuccess, maxInliers, pct,
    inlrNum, inlrNum_refined).  T is None where the reference returns [] (:75-89).
    A 3-point fit that the rank guard rejects scores 0 inliers (the reference would throw an
    uncaught error at :48 -- latent bug, not replicated)."""
    pts1 = np.asarray(pts1, dtype=np.float64)
    pts2 = np.asarray(pts2, dtype=np.float64)
    triplets = np.asarray(triplets, dtype=np.int64)
    iterNum = triplets.shape[0]
    thDist = float(coef["thDist"])
    ptNum = pts1.shape[0]
    thInlr = float(matlab_round(float(coef["thInlrRatio"]) * ptNum))      # :28
    REFINE = bool(coef.get("REFINE", True))
    inlrNum = np.zeros(iterNum, dtype=np.int64)
    inlrNum_ref = np.zeros(iterNum, dtype=np.int64)
    TForms = [None] * iterNum
    for p in range(iterNum):                                              # :40
        s = triplets[p]
        f1 = estimateTransform(pts1[s], pts2[s], reflection_fix)          # :45
        if f1 is None:
            continue
        dist = calcDists(f1, pts1, pts2)                                  # :48
        inl = np.nonzero(dist < thDist)[0]                                # :49
        inlrNum[p] = inl.size
        if inl.size >= thInlr:                                            # :53
            if REFINE:
                f1r = estimateTransform(pts1[inl], pts2[inl], reflection_fix)   # :55
                if f1r is None:
                    continue
                dist = calcDists(f1r, pts1, pts2)                         # :56
                inlrNum_ref[p] = int(np.sum(dist < thDist))               # :57-58
                if inlrNum_ref[p] >= thInlr:                              # :59-61
                    TForms[p] = f1r
            else:
                TForms[p] = f1                                            # :63
    counts = inlrNum_ref if REFINE else inlrNum
    idx = int(np.argmax(counts)) if iterNum else 0                        # :69-73 first arg-max
    T = TForms[idx] if iterNum else None
    if T is None:                                                         # :75-89
        return dict(T=None, inlierIdx=np.zeros(0, dtype=np.int64), numSuccess=0, maxInliers=0,
                    pct=0.0, inlrNum=inlrNum, inlrNum_refined=inlrNum_ref, best=-1)
    dist = calcDists(T, pts1, pts2)                                       # :78
    inlierIdx = np.nonzero(dist < thDist)[0]                              # :92
    numSuccess = int(np.sum(counts >= thInlr))                            # :94-98
    maxInliers = int(counts[idx])
    return dict(T=T, inlierIdx=inlierIdx, numSuccess=numSuccess, maxInliers=maxInliers,
                pct=100.0 * maxInliers / ptNum, inlrNum=inlrNum, inlrNum_refined=inlrNum_ref,
                best=idx)


def _splitmix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return x ^ (x >> np.uint64(31))


def ransac_triplets(seed: int, iterNum: int, P: int) -> np.ndarray:
    """The drop-in's documented stand-in for randperm(ptNum)(1:3) (ransac.m:42-43; MATLAB's RNG stream cannot
    be matched): restatement of include/pcreg.h pcreg_ransac_run.  Returns int32 [iterNum, 3], 0-based."""
    with np.errstate(over="ignore"):
        h = np.arange(iterNum, dtype=np.uint64)
        g = np.uint64(0x9E3779B97F4A7C15)
        s = np.uint64(seed & (2 ** 64 - 1))
        u = [_splitmix64((s + g * (np.uint64(3) * h + np.uint64(k + 1))).astype(np.uint64)) for k in range(3)]
    i0 = (u[0] % np.uint64(P)).astype(np.int64)
    i1 = (u[1] % np.uint64(P - 1)).astype(np.int64)
    i1 = i1 + (i1 >= i0)
    i2 = (u[2] % np.uint64(P - 2)).astype(np.int64)
    lo, hi = np.minimum(i0, i1), np.maximum(i0, i1)
    i2 = i2 + (i2 >= lo)
    i2 = i2 + (i2 >= hi)
    return np.stack([i0, i1, i2], axis=1).astype(np.int32)


def check_alignment(R1, R2) -> float:
    """checkAlignment, visualizeGTMatches.m:417-421 -- ||R1*R2' - I||_F (success < 0.5, :221)."""
    return float(np.linalg.norm(np.asarray(R1) @ np.asarray(R2).T - np.eye(3), "fro"))
