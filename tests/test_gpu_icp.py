"""Parity of the batched ICP (CUDA, through the C ABI) against the oracle composition
(oracle/icp.py).  Tolerances are BASELINE.json's: correspondence indices
bit-exact, rotation Frobenius <= 1e-6, translation <= 1e-6 relative, RMSE <= 1e-7 relative."""
import json
import os

import numpy as np
import pytest

import oracle
from pcreg_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

R_TOL = 1e-6
T_REL = 1e-6
RMSE_REL = 1e-7


def _compare(res, ref, check_idx=True):
    H = ref["T"].shape[0]
    for h in range(H):
        Rg, Ro = res["T"][h][:3, :3], ref["T"][h][:3, :3]
        tg, to = res["T"][h][3, :3], ref["T"][h][3, :3]
        assert np.linalg.norm(Rg - Ro, "fro") <= R_TOL, "hyp %d rotation off by %g" % (h, np.linalg.norm(Rg - Ro, "fro"))
        assert np.linalg.norm(tg - to) <= T_REL * max(np.linalg.norm(to), 1e-12), "hyp %d translation" % h
        if np.isnan(ref["rmse"][h]):
            assert np.isnan(res["rmse"][h])
        else:
            assert abs(res["rmse"][h] - ref["rmse"][h]) <= RMSE_REL * ref["rmse"][h], "hyp %d rmse %r vs %r" % (h, res["rmse"][h], ref["rmse"][h])
    assert np.array_equal(res["n_used"], ref["n_used"])
    assert np.array_equal(res["status"], ref["status"])
    # winner = first arg-min of rmse; hypotheses that converge to the same optimum have rmse equal to
    # rounding, so the index is only REQUIRED to match when the runner-up is clearly worse
    rr = np.asarray(ref["rmse"], dtype=np.float64)
    if ref["best"] < 0:
        assert res["best"] < 0
    else:
        others = np.delete(rr, ref["best"])
        others = others[~np.isnan(others)]
        if others.size == 0 or others.min() > rr[ref["best"]] * (1 + 1e-9):
            assert res["best"] == ref["best"]
        else:
            assert abs(rr[res["best"]] - rr[ref["best"]]) <= 1e-9 * rr[ref["best"]]
    if check_idx:
        assert np.array_equal(res["idx"], ref["idx"]), "%d correspondence mismatches" % int(np.sum(res["idx"] != ref["idx"]))


def _problem(nm, ns, H, seed, sigma=0.3, max_deg=8.0):
    model = synth.make_model(nm, seed)
    src, T_gt, c = synth.make_source(model, ns, sigma, seed + 1)
    T0 = synth.pose_grid(T_gt, c, max(1, H // 4), (2, 2, 1), max_deg, 1.0, seed + 2)[:H]
    return model, src, T0, T_gt


@pytest.mark.parametrize("nn", ["brute", "grid", "grid-novox"])
@pytest.mark.parametrize("mode", ["plain", "knn", "weighted"])
def test_icp_modes_match_oracle(pcreg, mode, nn):
    model, src, T0, _ = _problem(20_000, 700, 6, 100)
    g = synth.rng(77)
    w_src = g.uniform(0.5, 1.0, src.shape[0]) if mode == "weighted" else None
    omode = dict(plain=oracle.ICP_PLAIN, knn=oracle.ICP_KNN, weighted=oracle.ICP_WEIGHTED)[mode]
    gmode = dict(plain=pcreg.ICP_PLAIN, knn=pcreg.ICP_KNN, weighted=pcreg.ICP_WEIGHTED)[mode]
    thd = 9.0 if mode == "plain" else 0.0
    ref = oracle.icp_batch(model, src, T0, mode=omode, iters=12, k_frac=0.85, R_w=3.5, thDist2=thd, w_src=w_src, return_hist=True)
    m = pcreg.Model(model, grid=True, voxel_map=-1 if nn == "grid-novox" else 0)
    res = pcreg.icp_batch(m, src, T0, mode=gmode, iters=12, k_frac=0.85, R_w=3.5, thDist2=thd, w_src=w_src,
                          nn=pcreg.NN_BRUTE if nn == "brute" else pcreg.NN_GRID, return_idx=True, return_hist=True)
    _compare(res, ref)
    # per-iteration rmse history matches the oracle's trace too
    for h in range(T0.shape[0]):
        oh = ref["results"][h].get("rmse_hist")
        if oh is not None:
            np.testing.assert_allclose(res["rmse_hist"][h], oh, rtol=1e-7)
    m.destroy()


def test_icp_config1_knn_50_iterations(pcreg):
    """BASELINE.json configs[0]: 2k sparse cloud vs 50k model, known rigid TF + 1 mm noise,
    AlignPoints_KNN-style trimmed ICP, 50 iterations."""
    model = synth.make_model(50_000, 1001)
    src, T_gt, c = synth.make_source(model, 2000, 1.0, 1001)
    T0 = synth.perturb_pose(T_gt, c, synth.rot_axis_angle([0.3, -0.5, 0.8], np.deg2rad(5.0)), np.array([1.2, -1.0, 1.2]))[None]
    ref = oracle.icp_batch(model, src, T0, mode=oracle.ICP_KNN, iters=50, k_frac=0.85)
    m = pcreg.Model(model, grid=True)
    for nn in (pcreg.NN_BRUTE, pcreg.NN_GRID):
        res = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=50, k_frac=0.85, nn=nn, return_idx=True)
        _compare(res, ref)
    # and it actually registers: within a few tenths of a mm / degree of the ground truth
    assert oracle.check_alignment(res["T"][0][:3, :3], T_gt[:3, :3]) < 0.05
    m.destroy()


def test_icp_frozen_hypothesis_status(pcreg):
    """thDist2 so small that fewer than 3 correspondences survive: pose frozen, status 1, rmse NaN or tiny."""
    model, src, T0, _ = _problem(5000, 200, 3, 300)
    ref = oracle.icp_batch(model, src, T0, mode=oracle.ICP_PLAIN, iters=5, thDist2=1e-12)
    m = pcreg.Model(model, grid=True)
    res = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_PLAIN, iters=5, thDist2=1e-12, return_idx=True)
    assert np.array_equal(res["status"], ref["status"]) and res["status"].all()
    np.testing.assert_array_equal(res["T"], np.asarray(T0))
    assert res["best"] == ref["best"]
    m.destroy()


def test_icp_duplicate_source_points_tie_rule(pcreg):
    """Duplicated source points give exactly equal residuals at the trim boundary: the stable
    (lower index first) rule of MATLAB sort (AlignPoints_KNN.m:24) must be reproduced."""
    model, src, T0, _ = _problem(8000, 150, 2, 400)
    src = np.vstack([src, src, src[:37]])
    ref = oracle.icp_batch(model, src, T0, mode=oracle.ICP_KNN, iters=6, k_frac=0.5)
    m = pcreg.Model(model, grid=False)
    res = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=6, k_frac=0.5, return_idx=True)
    _compare(res, ref)
    m.destroy()


def test_icp_single_precision_source(pcreg):
    model, src, T0, _ = _problem(6000, 300, 2, 500)
    src32 = src.astype(np.float32)
    ref = oracle.icp_batch(model, src32.astype(np.float64), T0, mode=oracle.ICP_WEIGHTED, iters=8)
    m = pcreg.Model(model, grid=True)
    res = pcreg.icp_batch(m, src32, T0, mode=pcreg.ICP_WEIGHTED, iters=8, nn=pcreg.NN_GRID, return_idx=True)
    _compare(res, ref)
    m.destroy()


def test_icp_golden_fixture(pcreg):
    """Committed oracle output (tests/golden/make_golden.py) -- parity without running the oracle."""
    with open(os.path.join(HERE, "golden", "icp_small.json")) as f:
        G = json.load(f)
    model = synth.make_model(G["nm"], G["seed"])
    src, T_gt, c = synth.make_source(model, G["ns"], G["sigma"], G["seed"] + 1)
    T0 = np.asarray(G["T0"])
    m = pcreg.Model(model, grid=True)
    for mode_name, gm in (("plain", pcreg.ICP_PLAIN), ("knn", pcreg.ICP_KNN), ("weighted", pcreg.ICP_WEIGHTED)):
        want = G["modes"][mode_name]
        for nn in (pcreg.NN_BRUTE, pcreg.NN_GRID):
            res = pcreg.icp_batch(m, src, T0, mode=gm, iters=G["iters"], nn=nn, return_idx=True)
            ref = dict(T=np.asarray(want["T"]), rmse=np.asarray(want["rmse"]), n_used=np.asarray(want["n_used"], dtype=np.int32),
                       status=np.asarray(want["status"], dtype=np.int32), best=want["best"], idx=np.asarray(want["idx"], dtype=np.int32))
            _compare(res, ref)
    m.destroy()


def test_icp_larger_batch_properties(pcreg):
    """Size-independent properties at a size the oracle would not finish quickly: every hypothesis'
    rmse is finite, correspondences are valid indices, re-running is bit-identical (determinism),
    brute and grid agree bit for bit, and the winner is the first arg-min."""
    model = synth.make_model(200_000, 1003)
    src, T_gt, c = synth.make_source(model, 3000, 0.3, 1003)
    T0 = synth.pose_grid(T_gt, c, 4, (4, 4, 2), 10.0, 2.0, 5)
    m = pcreg.Model(model, grid=True)
    a = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=10, nn=pcreg.NN_GRID, return_idx=True)
    b = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=10, nn=pcreg.NN_GRID, return_idx=True)
    c2 = pcreg.icp_batch(m, src, T0[:16], mode=pcreg.ICP_KNN, iters=10, nn=pcreg.NN_BRUTE, return_idx=True)
    assert np.array_equal(a["T"], b["T"]) and np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["rmse"], b["rmse"])
    assert np.array_equal(a["T"][:16], c2["T"]) and np.array_equal(a["idx"][:16], c2["idx"])
    assert np.all(np.isfinite(a["rmse"])) and a["idx"].min() >= 0 and a["idx"].max() < model.shape[0]
    assert a["best"] == int(np.argmin(a["rmse"]))
    # rigid: R orthonormal to 1e-12
    R = a["T"][:, :3, :3]
    assert np.max(np.abs(R @ np.swapaxes(R, 1, 2) - np.eye(3))) < 1e-12
    m.destroy()


def test_icp_candidate_lists_are_exact_and_used(pcreg):
    """Grid NN answers most steady-state queries from per-query candidate lists (nn_grid.cu: k_nn_list).  A list
    answer is only accepted when it is provably the exact nearest neighbour, so 30 iterations of grid ICP must
    stay bit-identical to brute-force ICP (any differing correspondence would change the pose sums), and the
    profile must show that the list path really ran."""
    model = synth.make_model(150_000, 2024)
    src, T_gt, c = synth.make_source(model, 2500, 0.3, 2025)
    T0 = synth.pose_grid(T_gt, c, 4, (2, 2, 2), 8.0, 1.5, 9)
    m = pcreg.Model(model, grid=True, voxel_map=-1)          # per-query lists are the path of models WITHOUT a voxel map
    pcreg.set_profiling(True)
    a = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=30, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    prof = pcreg.last_profile()
    pcreg.set_profiling(False)
    b = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=30, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True)
    assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["T"], b["T"])
    assert np.array_equal(a["rmse_hist"], b["rmse_hist"])
    assert prof["certified_queries"] > 0.3 * prof["nn_queries"], prof
    m.destroy()


@pytest.mark.parametrize("vm", [-1, 1])
def test_icp_chunked_hypotheses_identical(pcreg, monkeypatch, vm):
    """Large batches are processed in chunks of hypotheses (bounded scratch: correspondences, candidate lists, the
    extension pool are per chunk).  Forcing 5 hypotheses per chunk must not change a single bit."""
    model = synth.make_model(60_000, 77)
    src, T_gt, c = synth.make_source(model, 1200, 0.3, 78)
    T0 = synth.pose_grid(T_gt, c, 4, (2, 2, 2), 8.0, 1.5, 9)[:23]
    m = pcreg.Model(model, grid=True, voxel_map=vm)
    a = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=12, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    monkeypatch.setenv("PCREG_MAX_CHUNK_HYP", "5")
    b = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=12, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    monkeypatch.delenv("PCREG_MAX_CHUNK_HYP")
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k], b[k]), k
    assert a["best"] == b["best"]
    m.destroy()


def test_icp_wide_balls_use_extension_lists(pcreg):
    """Source points far from the model (outliers the trim discards) have wide search balls: their candidate lists
    overflow the 64-entry row into the extension pool.  Grid ICP must still equal brute-force ICP bit for bit."""
    model = synth.make_model(150_000, 31)
    src, T_gt, c = synth.make_source(model, 1500, 0.3, 32)
    g = synth.rng(33)
    src = np.vstack([src, src[:300] + g.normal(0, 2.5, (300, 3))])          # 17 % gross outliers, 2-6 mm off the surface
    T0 = synth.pose_grid(T_gt, c, 2, (2, 2, 1), 5.0, 1.0, 9)
    m = pcreg.Model(model, grid=True, voxel_map=-1)
    pcreg.set_profiling(True)
    a = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=25, nn=pcreg.NN_GRID, return_idx=True)
    prof = pcreg.last_profile()
    pcreg.set_profiling(False)
    b = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=25, nn=pcreg.NN_BRUTE, return_idx=True)
    assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["T"], b["T"]) and np.array_equal(a["rmse"], b["rmse"])
    assert prof["certified_queries"] > 0.3 * prof["nn_queries"], prof
    # lazy trimming: most of the gross outliers are provably outside the 85 % trim from one pass to the next, their
    # searches are skipped (yet the final correspondences above are exact, and the poses identical to brute force)
    assert prof["lazy_skipped_queries"] > 0.05 * prof["nn_queries"], prof
    m.destroy()


@pytest.mark.parametrize("nn", ["brute", "grid", "grid-novox"])
def test_icp_two_lanes_identical(pcreg, monkeypatch, nn):
    """Large batches run as two sub-batches on two streams (icp.cu: lanes).  Forcing the two-lane schedule on a small
    batch -- with an odd hypothesis count and several chunks per lane -- must not change a single bit."""
    model = synth.make_model(60_000, 91)
    src, T_gt, c = synth.make_source(model, 1100, 0.3, 92)
    T0 = synth.pose_grid(T_gt, c, 4, (2, 2, 2), 8.0, 1.5, 9)[:27]
    m = pcreg.Model(model, grid=True, voxel_map=-1 if nn == "grid-novox" else 0)
    kind = pcreg.NN_GRID if nn.startswith("grid") else pcreg.NN_BRUTE
    monkeypatch.setenv("PCREG_LANES", "1")
    a = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=12, nn=kind, return_idx=True, return_hist=True)
    monkeypatch.setenv("PCREG_LANES", "2")
    b = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=12, nn=kind, return_idx=True, return_hist=True)
    monkeypatch.setenv("PCREG_MAX_CHUNK_HYP", "4")
    c2 = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=12, nn=kind, return_idx=True, return_hist=True)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c2[k]), k
    assert a["best"] == b["best"] == c2["best"]
    m.destroy()


@pytest.mark.parametrize("mode", ["knn", "weighted", "reject"])
@pytest.mark.parametrize("cpp", [0.0, 0.08])
def test_icp_warp_row_scan_identical(pcreg, monkeypatch, mode, cpp):
    """nn_grid.cu has two row-scan kernels and two pyramid-walk kernels: per-lane (sparse models) and warp-per-query with
    coalesced runs (dense models, picked by points per occupied cell; the warp walk hands queries whose frontier outgrows its
    buffer to the per-lane walk).  All of them, and the brute-force path, must return the same bits -- on the default grid and on a deliberately coarse one (cells_per_point 0.08: ~15-40 points per occupied cell,
    so the dense path is also what the launcher picks by itself, and the candidate lists overflow into extension slots)."""
    model = synth.make_model(80_000, 191)
    src, T_gt, c = synth.make_source(model, 1500, 0.3, 192)
    T0 = synth.pose_grid(T_gt, c, 3, (2, 2, 2), 8.0, 1.5, 19)[:21]
    m = pcreg.Model(model, grid=True, cells_per_point=cpp, voxel_map=-1)
    kw = dict(knn=dict(mode=pcreg.ICP_KNN), weighted=dict(mode=pcreg.ICP_WEIGHTED, R_w=3.5),
              reject=dict(mode=pcreg.ICP_PLAIN, thDist2=4.0))[mode]
    ref = pcreg.icp_batch(m, src, T0, iters=14, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True, **kw)
    out = {}
    for rs in ("lane", "warp", "warp-overflow", None):
        for var in ("PCREG_ROWSCAN", "PCREG_WALK", "PCREG_WW_CAP"):
            monkeypatch.delenv(var, raising=False)
        if rs is not None:                                   # the same choice for the row scan and for the pyramid walk
            monkeypatch.setenv("PCREG_ROWSCAN", rs.split("-")[0])
            monkeypatch.setenv("PCREG_WALK", rs.split("-")[0])
        if rs == "warp-overflow":                            # a frontier of 8 nodes: most walked queries are handed on to the per-lane walk
            monkeypatch.setenv("PCREG_WW_CAP", "8")
        out[rs] = pcreg.icp_batch(m, src, T0, iters=14, nn=pcreg.NN_GRID, return_idx=True, return_hist=True, **kw)
    for rs, r in out.items():
        for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
            assert np.array_equal(r[k], ref[k], equal_nan=True) if r[k].dtype.kind == "f" else np.array_equal(r[k], ref[k]), (rs, k)
        assert r["best"] == ref["best"]
    m.destroy()
