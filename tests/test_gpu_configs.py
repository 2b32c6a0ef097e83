"""BASELINE.json configs[3] and configs[4] as parity cases at their full MODEL sizes (2 M and 16 M points, 20 k and 65 k
source points) with a reduced number of hypotheses: the oracle checks what it can finish in seconds, the rest is held
by size-independent properties -- the two independent exact searches (FP32-scan brute force with FP64 decision, grid +
pyramid + candidate lists) must agree bit for bit on indices, distances, poses and RMSE."""
import numpy as np
import pytest

import oracle
from pcreg_b200 import synth
from test_gpu_icp import _compare

pytestmark = pytest.mark.gpu


def test_config4_ransac_seeded_polish_2m_model(pcreg):
    """C4: hypotheses as a RANSAC stage seeds them (true pose perturbed by a few degrees / mm), 20-iteration ICP polish,
    20 k source vs 2 M model, grid NN, PLAIN mode with the squared rejection threshold thDist2 = 4 (ransac.m:49)."""
    model = synth.make_model(2_000_000, 1004)
    src, T_gt, c = synth.make_source(model, 20_000, 0.3, 1004)
    g = synth.rng(5)
    H = 48
    T0 = np.stack([synth.perturb_pose(T_gt, c, synth.rot_axis_angle(g.standard_normal(3), np.deg2rad(g.uniform(0, 6))),
                                      g.normal(0, 1.0, 3)) for _ in range(H)])
    m = pcreg.Model(model, grid=True)
    res = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_PLAIN, iters=20, thDist2=4.0, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    bru = pcreg.icp_batch(m, src, T0[:3], mode=pcreg.ICP_PLAIN, iters=20, thDist2=4.0, nn=pcreg.NN_BRUTE, return_idx=True)
    m.destroy()
    assert np.array_equal(res["idx"][:3], bru["idx"]) and np.array_equal(res["T"][:3], bru["T"])
    assert np.array_equal(res["rmse"][:3], bru["rmse"]) and np.array_equal(res["n_used"][:3], bru["n_used"])
    # the oracle composition (FP64 numpy + exact-ified kd-tree) on two of the hypotheses
    ref = oracle.icp_batch(model, src, T0[:2], mode=oracle.ICP_PLAIN, iters=20, thDist2=4.0)
    sub = {k: (v[:2] if isinstance(v, np.ndarray) else v) for k, v in res.items()}
    sub["best"] = int(np.argmin(sub["rmse"]))
    _compare(sub, ref)
    # and the polish does its job (point-to-point ICP slides slowly along the surface: 20 iterations roughly halve the
    # seed error, the residual drops to the 0.3 mm noise floor; values cross-checked with the oracle on all 48)
    errs0 = np.array([oracle.check_alignment(T0[h][:3, :3], T_gt[:3, :3]) for h in range(H)])
    errs = np.array([oracle.check_alignment(res["T"][h][:3, :3], T_gt[:3, :3]) for h in range(H)])
    assert np.all(res["status"] == 0) and np.median(errs) < 0.75 * np.median(errs0) and errs.max() < 0.2
    assert np.all(res["rmse_hist"][:, -1] < res["rmse_hist"][:, 0]) and res["rmse"].max() < 0.4
    assert np.all(res["n_used"] > 0.9 * src.shape[0])                 # thDist2 = 4 rejects only far starts


def test_config5_16m_model_grid_search(pcreg):
    """C5: 16 M-point model (the upsampleMesh.m-sized cloud), 65 536 source points, grid NN.  One search of the whole
    source: grid == brute bit for bit, and == the oracle's FP64 brute force on a sample; then a small multi-start ICP
    (KNN trim, 20 iterations) whose grid path must reproduce the brute-force path exactly."""
    model = synth.make_model(16_000_000, 1005)
    src, T_gt, c = synth.make_source(model[::16], 65_536, 0.3, 1005)
    q = synth.apply_T(src, T_gt)
    m = pcreg.Model(model, grid=True)
    info = m.grid_info()
    assert info["occupied"] > 1_000_000 and min(info["dims"]) > 100
    gi, gd = m.nn_search(q, pcreg.NN_GRID)
    bi, bd = m.nn_search(q, pcreg.NN_BRUTE)
    assert np.array_equal(gi, bi) and np.array_equal(gd, bd)
    sel = np.arange(0, q.shape[0], 341)[:192]
    oi, od = oracle.nn_brute(model, q[sel])
    assert np.array_equal(gi[sel], oi) and np.array_equal(gd[sel], od)
    T0 = synth.pose_grid(T_gt, c, 2, (2, 2, 1), 10.0, 2.0, 7)           # 8 hypotheses
    res = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=20, nn=pcreg.NN_GRID, return_idx=True)
    bru = pcreg.icp_batch(m, src, T0[:2], mode=pcreg.ICP_KNN, iters=20, nn=pcreg.NN_BRUTE, return_idx=True)
    m.destroy()
    assert np.array_equal(res["idx"][:2], bru["idx"]) and np.array_equal(res["T"][:2], bru["T"]) and np.array_equal(res["rmse"][:2], bru["rmse"])
    assert np.all(res["n_used"] == oracle.matlab_round(0.85 * 65_536))   # AlignPoints_KNN.m:20-21
    # values cross-checked with the oracle composition (kd-tree, 76 s on 8 cores): alignment error 0.028-0.058, rmse 0.17-0.19
    assert oracle.check_alignment(res["T"][res["best"]][:3, :3], T_gt[:3, :3]) < 0.06 and res["rmse"].max() < 0.21


def _edge_hypotheses(w):
    """indices into bench.make_inputs' pose grid (rotation-major, then x, y, z): corners of the translation lattice under
    the first, a middle and the last rotations -- the hypotheses that start farthest from the optimum"""
    nx, ny, nz = w["trans"]
    out = []
    for r, (x, y, z) in zip((w["rot"] - 1, w["rot"] - 1, w["rot"] // 2, w["rot"] // 2, 1, 1, w["rot"] - 2, 2),
                            ((0, 0, 0), (nx - 1, ny - 1, nz - 1), (0, ny - 1, 0), (nx - 1, 0, nz - 1), (0, 0, nz - 1), (nx - 1, ny - 1, 0),
                             (0, ny - 1, nz - 1), (nx - 1, 0, 0))):
        out.append(((r * nx + x) * ny + y) * nz + z)
    return np.array(out)


def test_config3_full_size_vs_oracle_and_brute(pcreg):
    """BASELINE.json configs[2] exactly as bench.py runs it (1 M-point model, 5 000 source points, 4096-pose grid of 20 deg /
    2 mm, KNN trim, 30 iterations, seed 1003): eight hypotheses from the EDGES of the pose grid against the oracle
    composition, and 128 hypotheses spread over the grid grid-path == brute-force path bit for bit."""
    import bench
    w = bench.WORKLOADS["c3"]
    model, src, T0, w_src, T_gt = bench.make_inputs(w, 0)
    assert model.shape[0] == 1_000_000 and src.shape[0] == 5000 and T0.shape[0] == 4096
    sel = _edge_hypotheses(w)
    ref = oracle.icp_batch(model, src, T0[sel], mode=oracle.ICP_KNN, iters=30, k_frac=0.85, return_hist=True)
    m = pcreg.Model(model, grid=True)
    assert m.voxel_info()["voxels"] > 0                      # the path bench.py measures: Voronoi voxel map + fused kernel
    res = pcreg.icp_batch(m, src, T0[sel], mode=pcreg.ICP_KNN, iters=30, k_frac=0.85, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    _compare(res, ref)
    for h in range(sel.size):
        np.testing.assert_allclose(res["rmse_hist"][h], ref["results"][h]["rmse_hist"], rtol=1e-7)
    many = np.arange(0, 4096, 32)                            # 128 hypotheses over the whole grid
    a = pcreg.icp_batch(m, src, T0[many], mode=pcreg.ICP_KNN, iters=30, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    b = pcreg.icp_batch(m, src, T0[many], mode=pcreg.ICP_KNN, iters=30, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k], b[k]), k
    # the same batch through the grid kernels of models WITHOUT a voxel map (pyramid walk, row scan, candidate lists)
    m2 = pcreg.Model(model, grid=True, voxel_map=-1)
    c2 = pcreg.icp_batch(m2, src, T0[many[:32]], mode=pcreg.ICP_KNN, iters=30, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(c2[k], b[k][:32]), k
    m.destroy(); m2.destroy()


def test_config2_full_size_vs_oracle(pcreg):
    """BASELINE.json configs[1] as bench.py runs it: one AlignPoints_weighted-style alignment, 10 000 weighted source points
    vs a 500 k-point model, 100 iterations -- brute force and grid NN against the oracle composition."""
    import bench
    w = bench.WORKLOADS["c2"]
    model, src, T0, w_src, T_gt = bench.make_inputs(w, 0)
    assert model.shape[0] == 500_000 and src.shape[0] == 10_000 and T0.shape[0] == 1 and w_src is not None
    ref = oracle.icp_batch(model, src, T0, mode=oracle.ICP_WEIGHTED, iters=100, R_w=3.5, w_src=w_src, return_hist=True)
    m = pcreg.Model(model, grid=True)
    for nn in (pcreg.NN_BRUTE, pcreg.NN_GRID):
        res = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_WEIGHTED, iters=100, R_w=3.5, w_src=w_src, nn=nn, return_idx=True, return_hist=True)
        _compare(res, ref)
        np.testing.assert_allclose(res["rmse_hist"][0], ref["results"][0]["rmse_hist"], rtol=1e-7)
    m.destroy()


def test_config5_bench_shape_wide_starts(pcreg):
    """C5 at the shape bench.py --workload c5 runs (16 M-point model, 65 536 source points, the 10 deg / 2 mm pose grid):
    64 hypotheses spread over the grid, including its edges -- wide search balls that exercise the warp-per-query walk, its
    overflow hand-off and the extension pool of the candidate lists.  Grid path == brute-force path bit for bit on the four
    farthest starts; determinism and the trim count on all 64."""
    import bench
    w = bench.WORKLOADS["c5"]
    model, src, T0, w_src, T_gt = bench.make_inputs(w, 0)
    assert model.shape[0] == 16_000_000 and src.shape[0] == 65_536
    sel = np.unique(np.concatenate([_edge_hypotheses(w), np.arange(0, T0.shape[0], T0.shape[0] // 56)]))[:64]
    m = pcreg.Model(model, grid=True)
    a = pcreg.icp_batch(m, src, T0[sel], mode=pcreg.ICP_KNN, iters=20, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    a2 = pcreg.icp_batch(m, src, T0[sel], mode=pcreg.ICP_KNN, iters=20, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k], a2[k]), k
    far = np.argsort(-a["rmse_hist"][:, 0])[:4]              # the four starts with the largest initial residual
    b = pcreg.icp_batch(m, src, T0[sel][far], mode=pcreg.ICP_KNN, iters=20, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k][far], b[k]), k
    assert np.all(a["n_used"] == oracle.matlab_round(0.85 * 65_536)) and np.all(np.isfinite(a["rmse"]))
    m.destroy()
