"""Generates the committed golden fixtures in tests/golden/ (run from the repo root:
`python tests/golden/make_golden.py`).

kat_estimateTransform.json -- the reference's ONE self-contained known-answer vector,
  testTransformEstimation.m:2-14: four hard-coded points, r = [0.1 0.2 0.3] through eul2rotm's default
  'ZYX' order, t = [1 2 3], pts_tf = pts*R + t.  Expected output (analytic):
  estimateTransform(pts_tf, pts) = [R 0; t 1] and [pts,1]*T = pts_tf.
kat_identities.json -- the analytic structure of testRANSAC.m:17-29,40-42 (T = [R 0; t 1],
  T_back = [R' 0; -t*R' 1]) for r = [1.5 -1.2 0.8], t = [1 2 3].
icp_small.json -- outputs of the ORACLE composition (oracle/icp.py) on a small seeded problem, so the
  GPU parity test can run without the oracle present.
rows_small.json -- inputs AND oracle outputs of the other rows of the path on small seeded problems: the six AlignPoints*
  functions (ten call variants) on one neighbourhood, the whole ransac.m call on the documented seeded samples,
  getLocalPoints for four centres (one of them empty), getMatches on count-like descriptors whose every decision
  has a margin above 1e-6, and getSpacialHistogramDescriptors (16 keypoints, 8 survive the point-count / variance checks).  tests/test_oracle.py checks that the oracle still reproduces it, tests/test_gpu_golden_rows.py
  checks the CUDA path against it.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pcreg_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    pts = np.array([[1, 5, 7], [4, 9, 3], [9, 3, 4], [1, 2, 4]], dtype=np.float64)   # testTransformEstimation.m:2-5
    r = [0.1, 0.2, 0.3]                                                             # :8
    t = np.array([1.0, 2.0, 3.0])                                                    # :9
    R = oracle.eul2rotm(r)                                                           # :12 (default 'ZYX')
    pts_tf = pts @ R + t                                                             # :14
    T = np.eye(4)
    T[:3, :3] = R
    T[3, :3] = t
    with open(os.path.join(HERE, "kat_estimateTransform.json"), "w") as f:
        json.dump(dict(source="testTransformEstimation.m:2-14", pts=pts.tolist(), r=r, t=t.tolist(),
                       R=R.tolist(), pts_tf=pts_tf.tolist(), T=T.tolist()), f, indent=1)

    r2 = [1.5, -1.2, 0.8]                                                            # testRANSAC.m:17-18
    R2 = oracle.eul2rotm(r2)
    T2 = np.eye(4); T2[:3, :3] = R2; T2[3, :3] = t
    T2b = np.eye(4); T2b[:3, :3] = R2.T; T2b[3, :3] = -t @ R2.T                      # testRANSAC.m:40-42
    with open(os.path.join(HERE, "kat_identities.json"), "w") as f:
        json.dump(dict(source="testRANSAC.m:17-29,40-42", r=r2, t=t.tolist(), T_true=T2.tolist(), T_back=T2b.tolist()), f, indent=1)

    nm, ns, seed, sigma, iters = 6000, 250, 4242, 0.3, 8
    model = synth.make_model(nm, seed)
    src, T_gt, c = synth.make_source(model, ns, sigma, seed + 1)
    T0 = synth.pose_grid(T_gt, c, 2, (2, 1, 1), 6.0, 1.0, seed + 2)
    g = dict(nm=nm, ns=ns, seed=seed, sigma=sigma, iters=iters, T0=T0.tolist(), modes={})
    for name, mode in (("plain", oracle.ICP_PLAIN), ("knn", oracle.ICP_KNN), ("weighted", oracle.ICP_WEIGHTED)):
        res = oracle.icp_batch(model, src, T0, mode=mode, iters=iters, brute=True)
        g["modes"][name] = dict(T=res["T"].tolist(), rmse=res["rmse"].tolist(), n_used=res["n_used"].tolist(),
                                status=res["status"].tolist(), best=res["best"], idx=res["idx"].tolist())
    with open(os.path.join(HERE, "icp_small.json"), "w") as f:
        json.dump(g, f)
    write_rows_small()
    print("golden fixtures written to", HERE)


ALIGN_CASES = [
    ("AlignPoints", lambda o, p: o.AlignPoints(p)),
    ("AlignPoints_KNN", lambda o, p: o.AlignPoints_KNN(p)),
    ("AlignPoints_KNN_C1", lambda o, p: o.AlignPoints_KNN(p, True, False)),
    ("AlignPoints_KNN_C2", lambda o, p: o.AlignPoints_KNN(p, False, True)),
    ("AlignPoints_KNN_C1C2", lambda o, p: o.AlignPoints_KNN(p, True, True)),
    ("AlignPoints_knn_60", lambda o, p: o.AlignPoints_knn(p, 60)),
    ("AlignPoints_knn_5000", lambda o, p: o.AlignPoints_knn(p, 5000)),
    ("AlignPoints_weighted", lambda o, p: o.AlignPoints_weighted(p)),
    ("AlignPoints_c", lambda o, p: o.AlignPoints_c(p)[:2]),
    ("AlignPoints_KNN_c", lambda o, p: o.AlignPoints_KNN_c(p)),
]
MATCH_PAR = dict(UNNORMALIZE=True, norm_factor=2, CHANGE_METRIC=True, metric_factor=0.6, Method="Approximate",
                 MatchThreshold=10, MaxRatio=0.99, Metric="SAD", Unique=True)          # completeExperiment.m:112-122
RANSAC_COEF = dict(thDist=0.25, thInlrRatio=0.1, REFINE=True, iterNum=400)


def rows_small(o):
    """Inputs (seeded) and the outputs of implementation `o` (the oracle package here; anything with the same functions)."""
    out = {}
    p = synth.make_neighbourhoods(1, 77, nmin=200, nmax=240)[0]
    al = {}
    for name, fn in ALIGN_CASES:
        r = fn(o, p)
        al[name] = dict(aligned=None if r[0] is None else np.asarray(r[0]).tolist(),
                        coeff=None if r[1] is None else np.asarray(r[1]).tolist(),
                        c=np.asarray(r[2]).tolist() if len(r) > 2 and r[2] is not None else None)
    out["align"] = dict(pts=p.tolist(), cases=al)

    p1, p2, _ = synth.make_ransac_problem(80, 0.4, 0.1, 78)
    seed = 2025
    tri = oracle.ransac_triplets(seed, RANSAC_COEF["iterNum"], p1.shape[0])
    r = o.ransac(p1, p2, RANSAC_COEF, tri)
    out["ransac"] = dict(p1=p1.tolist(), p2=p2.tolist(), seed=seed, coef=RANSAC_COEF, triplets_head=tri[:5].tolist(),
                         T=r["T"].tolist(), inlierIdx=np.asarray(r["inlierIdx"]).tolist(), numSuccess=int(r["numSuccess"]),
                         maxInliers=int(r["maxInliers"]), best=int(r["best"]))

    g = synth.rng(79)
    cloud = g.uniform(0, 10, (600, 3))
    centres = np.array([[5.0, 5.0, 5.0], [1.0, 9.0, 2.0], [7.5, 2.5, 6.0], [40.0, 40.0, 40.0]])
    lp = []
    for c in centres:
        q, d = o.getLocalPoints(cloud, 2.0, c, 5, 200)
        lp.append(dict(pts=None if q is None else np.asarray(q).tolist(), dists=None if d is None else np.asarray(d).tolist()))
    out["local_points"] = dict(cloud=cloud.tolist(), centres=centres.tolist(), R=2.0, min_points=5, max_points=200, results=lp)

    g = np.random.default_rng(80)
    base = g.gamma(0.6, 4.0, (14, 48))
    dS = g.poisson(base).astype(np.float64)
    dM = np.vstack([g.poisson(base * (1.0 + 0.1 * g.standard_normal(base.shape)).clip(0.0, None)),
                    g.poisson(g.gamma(0.6, 4.0, (16, 48)))]).astype(np.float64)
    pairs, metric = o.getMatches(dS, dM, MATCH_PAR, return_metric=True)
    out["matches"] = dict(descSurface=dS.tolist(), descModel=dM.tolist(), par=MATCH_PAR, pairs=np.asarray(pairs).tolist(),
                          metric=np.asarray(metric).tolist())
    # getSpacialHistogramDescriptors: the model is regenerated from its seed (as icp_small.json does), keypoints stored
    dmodel = np.asarray(synth.make_model(DESC_MODEL[0], DESC_MODEL[1]), dtype=np.float64)
    g = synth.rng(82)
    kp = np.vstack([dmodel[g.integers(0, dmodel.shape[0], 12)] + g.normal(0, 0.3, (12, 3)), g.uniform(dmodel.min(0), dmodel.max(0), (4, 3))])
    feat, desc = o.getSpacialHistogramDescriptors(dmodel, kp, DESC_OPTS) if o is oracle else (None, None)
    if feat is not None:
        out["descriptors"] = dict(model=list(DESC_MODEL), keypoints=kp.tolist(), opts=dict(DESC_OPTS, thVar=list(DESC_OPTS["thVar"])),
                                  feat=np.asarray(feat).tolist(), desc=np.asarray(desc).astype(int).tolist())
    return out


DESC_MODEL = (40_000, 81)            # synth.make_model(n, seed)
DESC_OPTS = dict(min_pts=40, max_pts=2500, R=3.5, thVar=(1.2, 1.5), k=0.85, ALIGN_POINTS=True)


def write_rows_small():
    out = rows_small(oracle)
    # the fixture must not sit on a decision boundary: every getMatches decision of the oracle has a margin > 1e-6
    m = out["matches"]
    _, _, info = oracle.getMatches(np.asarray(m["descSurface"]), np.asarray(m["descModel"]), MATCH_PAR, return_all=True)
    d1, d2, S, j1 = info["d1"], info["d2"], info["S"], info["j1"]
    assert np.all(np.abs(d2 - d1) > 1e-6 * d2) and np.all(np.abs(d1 - info["thr"]) > 1e-6 * info["thr"])
    assert np.all(np.abs(info["ratio"] - MATCH_PAR["MaxRatio"]) > 1e-6)
    srt = np.sort(S[:, j1], axis=0)
    assert np.all(srt[1] - srt[0] > 1e-6 * srt[1])
    assert len(m["pairs"]) >= 5
    assert out["ransac"]["maxInliers"] >= 20 and out["align"]["cases"]["AlignPoints_c"]["coeff"] is not None
    assert sum(r["pts"] is None for r in out["local_points"]["results"]) == 1
    assert 4 <= len(out["descriptors"]["feat"]) < 16
    with open(os.path.join(HERE, "rows_small.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
