"""The whole registration chain of the reference's drivers (slideMatchingWindow_v2.m:182-226, completeExperiment.m:
keypoints -> getSpacialHistogramDescriptors -> getMatches -> ransac -> pose) followed by the ICP polish, every stage on
the GPU through the public API, on one synthetic scene with a known ground-truth pose.  This is the "a user of the
reference can switch" check: the stages' outputs feed each other exactly as the .m drivers wire them."""
import numpy as np
import pytest

from pcreg_b200 import synth

pytestmark = pytest.mark.gpu

DESC_OPTS = dict(min_pts=500, max_pts=6000, R=3.5, thVar=(1.0, 1.0), k=0.85, ALIGN_POINTS=True)   # GetSphericalDescriptors.m:133-139
MATCH_PAR = dict(UNNORMALIZE=True, norm_factor=2, CHANGE_METRIC=True, metric_factor=0.6, Method="Approximate",
                 MatchThreshold=10, MaxRatio=0.99, Metric="SAD", Unique=True)                      # completeExperiment.m:112-122
RANSAC_COEF = dict(minPtNum=3, iterNum=20000, thDist=0.2, thInlrRatio=0.1, REFINE=True)          # slideMatchingWindow_v2.m:162-167


def make_scene(nm=300_000, seed=77, crop=14.0, n_kp_model=500, n_kp_surface=150, noise=0.02):
    """Model cloud, a cropped 'surface' cloud (the crop moved by the inverse of T_gt, slightly noisy), and keypoints:
    the model's inside the crop with their full neighbourhood (R_crop - 3.5, slideMatchingWindow_v2.m:182), the
    surface's = moved copies of some of them plus a small offset (independent detections of the same spots)."""
    g = synth.rng(seed)
    model = np.asarray(synth.make_model(nm, seed), dtype=np.float64)
    c = model[g.integers(0, nm)]
    d = np.linalg.norm(model - c, axis=1)
    T_gt = synth.make_T(synth.rot_xyz(g.uniform(0, 2 * np.pi, 3)), np.array([13.0, 25.0, -17.0]))    # slideMatchingWindow.m:35-36
    patch = model[d < crop]
    surface = synth.apply_T(patch + g.normal(0, noise, patch.shape), synth.invert_T(T_gt))
    inner = np.nonzero(d < crop - 3.5)[0]
    kpM = model[g.choice(inner, n_kp_model, replace=False)]
    kpS = synth.apply_T(kpM[:n_kp_surface] + g.normal(0, 0.05, (n_kp_surface, 3)), synth.invert_T(T_gt))
    return model, surface, kpM, kpS, T_gt


def run_chain(api, scene):
    model, surface, kpM, kpS, T_gt = scene
    mM, mS = api.Model(model), api.Model(surface)
    try:
        featM, descM = api.getSpacialHistogramDescriptors(mM, kpM, DESC_OPTS)
        featS, descS = api.getSpacialHistogramDescriptors(mS, kpS, DESC_OPTS)
    finally:
        mM.destroy(); mS.destroy()
    matches = api.getMatches(descS, descM, MATCH_PAR)
    loc1M, loc1S = featM[matches[:, 1]], featS[matches[:, 0]]                   # completeExperiment.m: matched locations
    res = api.ransac_seeded(loc1M, loc1S, RANSAC_COEF, seed=5)                  # [loc1S,1]*T = [loc1M,1]
    return featM, featS, matches, loc1M, loc1S, res


def test_descriptor_match_ransac_icp_chain(pcreg):
    scene = make_scene()
    model, surface, kpM, kpS, T_gt = scene
    featM, featS, matches, loc1M, loc1S, res = run_chain(pcreg, scene)
    assert featM.shape[0] > 300 and featS.shape[0] > 100
    assert matches.shape[0] > 50, "only %d matches" % matches.shape[0]          # the reference's gate (slideMatchingWindow_v2.m:170,196)
    true_pos = np.linalg.norm(synth.apply_T(loc1S, T_gt) - loc1M, axis=1) < 0.3
    assert true_pos.mean() > 0.5, "precision %.2f" % true_pos.mean()
    T = res["T"]
    assert T is not None and res["maxInliers"] >= 0.5 * matches.shape[0]
    # visualizeGTMatches.m:417-421 metric; success threshold 0.5 at :221
    assert np.linalg.norm(T[:3, :3] @ T_gt[:3, :3].T - np.eye(3)) < 0.05
    assert np.linalg.norm(synth.apply_T(surface, T) - synth.apply_T(surface, T_gt), axis=1).max() < 0.3
    # ICP polish of the RANSAC pose against the dense model (grid NN), 85 % trim
    g = synth.rng(9)
    src = surface[g.choice(surface.shape[0], 4000, replace=False)]
    m = pcreg.Model(model, grid=True)
    out = pcreg.icp_batch(m, src, T[None], mode=pcreg.ICP_KNN, iters=20, nn=pcreg.NN_GRID)
    m.destroy()
    Tp = out["T"][0]
    err_ransac = np.linalg.norm(synth.apply_T(surface, T) - synth.apply_T(surface, T_gt), axis=1).mean()
    err_icp = np.linalg.norm(synth.apply_T(surface, Tp) - synth.apply_T(surface, T_gt), axis=1).mean()
    assert err_icp < 0.05 and err_icp <= err_ransac + 0.01, (err_ransac, err_icp)
    assert out["rmse"][0] < 0.06                                                 # the noise floor (sigma 0.02 per axis) plus sampling
