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
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
