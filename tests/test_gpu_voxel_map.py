"""The Voronoi voxel map (pcreg_b200/csrc/nn_vox.cu) against the FP64 oracle and against the brute-force kernel.
Bar: indices and squared distances BIT-EXACT, every exact tie resolved to the smallest index -- whatever the voxel size,
the list cap, the margin, and for queries that fall back to the pyramid walk (voxel without a list, outside the box)."""
import numpy as np
import pytest

import oracle
from pcreg_b200 import synth

pytestmark = pytest.mark.gpu


def _nn_equal(pcreg, m, model, q):
    oi, od = oracle.nn_brute(model, q)
    pcreg.set_profiling(True)
    gi, gd = m.nn_search(q, pcreg.NN_GRID)
    prof = pcreg.last_profile()
    pcreg.set_profiling(False)
    bad = np.nonzero(gi != oi)[0]
    assert bad.size == 0, "%d index mismatches, first at %s: got %s want %s (d2 %s vs %s)" % (
        bad.size, bad[:5], gi[bad[:5]], oi[bad[:5]], gd[bad[:5]], od[bad[:5]])
    assert np.array_equal(gd, od)
    return prof


def test_voxel_map_is_built_and_answers_the_queries(pcreg):
    model = synth.make_model(200_000, 1003)
    src, T_gt, c = synth.make_source(model, 4000, 0.3, 1003)
    m = pcreg.Model(model, grid=True)
    info = m.voxel_info()
    assert info["voxels"] > 0 and info["entries"] > 0 and info["no_room"] == 0, info
    assert info["listed"] > 0.9 * info["voxels"], info
    assert info["entries"] / info["listed"] < 16, info           # short lists are the point of the structure
    # queries on the surface (converged ICP) and a few mm off it (first pass of a multi-start)
    g = synth.rng(4)
    q = np.vstack([synth.apply_T(src, T_gt), synth.apply_T(src, T_gt) + g.normal(0, 2.0, (4000, 3))])
    prof = _nn_equal(pcreg, m, model, q)
    assert prof["voxel_map"] == 1
    assert prof["certified_queries"] > 0.95 * q.shape[0], prof    # answered by the list scan, not by the walk
    m.destroy()


@pytest.mark.parametrize("scale", [0.6, 1.25, 3.0])
def test_voxel_map_any_voxel_size(pcreg, scale):
    model = synth.make_model(60_000, 21)
    g = synth.rng(22)
    q = np.vstack([np.asarray(model[:2000], dtype=np.float64) + g.normal(0, 0.5, (2000, 3)), g.uniform(-20, 120, (2000, 3))])
    m = pcreg.Model(model, grid=True, voxel_map=1, voxel_scale=scale)
    _nn_equal(pcreg, m, model, q)
    m.destroy()


def test_voxel_map_small_cap_falls_back_to_the_walk(pcreg, monkeypatch):
    """PCREG_VOX_CAP=8: most lists near the surface are over the cap and dropped; those queries are walked."""
    model = synth.make_model(60_000, 23)
    g = synth.rng(24)
    q = np.asarray(model[:3000], dtype=np.float64) + g.normal(0, 0.3, (3000, 3))
    monkeypatch.setenv("PCREG_VOX_CAP", "8")
    m = pcreg.Model(model, grid=True, voxel_map=1)
    monkeypatch.delenv("PCREG_VOX_CAP")
    info = m.voxel_info()
    assert info["too_long"] > 0, info
    prof = _nn_equal(pcreg, m, model, q)
    assert prof["walked_queries"] > 0, prof
    m.destroy()


def test_voxel_map_no_margin_queries_outside(pcreg):
    model = synth.make_model(40_000, 25)
    g = synth.rng(26)
    q = g.uniform(-60, 160, (4000, 3))                           # mostly outside the bounding box
    m = pcreg.Model(model, grid=True, voxel_map=1, voxel_margin=-1.0)
    prof = _nn_equal(pcreg, m, model, q)
    assert prof["walked_queries"] > 1000, prof
    m.destroy()


def test_voxel_map_few_voxels(pcreg):
    """max_voxels far below the automatic size: coarse voxels, long lists, many over the cap -- still exact."""
    model = synth.make_model(50_000, 27)
    g = synth.rng(28)
    q = np.asarray(model[:2500], dtype=np.float64) + g.normal(0, 1.0, (2500, 3))
    m = pcreg.Model(model, grid=True, voxel_map=1, max_voxels=30_000)
    assert 0 < m.voxel_info()["voxels"] <= 30_000
    _nn_equal(pcreg, m, model, q)
    m.destroy()


def test_voxel_map_lattice_ties(pcreg):
    ax = np.arange(10, dtype=np.float64)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    pts = np.column_stack([X.ravel(), Y.ravel(), Z.ravel()])
    g = synth.rng(5)
    model = np.vstack([pts, pts])[g.permutation(2 * pts.shape[0])]
    q = np.vstack([pts[:500] + 0.5, pts[:300], pts[:300] + np.array([0.5, 0.0, 0.0]), pts[:300] + np.array([0.5, 0.5, 0.0])])
    for scale in (0.5, 1.25):
        m = pcreg.Model(model, grid=True, voxel_map=1, voxel_scale=scale)
        _nn_equal(pcreg, m, model, q)
        m.destroy()


def test_voxel_map_icp_equals_brute_force_icp(pcreg):
    """30 iterations of multi-start ICP on the voxel map stay bit-identical to brute-force ICP (poses, RMSE history,
    final correspondences): any differing correspondence would change the pose sums."""
    model = synth.make_model(150_000, 2024)
    src, T_gt, c = synth.make_source(model, 2500, 0.3, 2025)
    T0 = synth.pose_grid(T_gt, c, 4, (2, 2, 2), 15.0, 2.0, 9)
    m = pcreg.Model(model, grid=True)
    assert m.voxel_info()["voxels"] > 0
    pcreg.set_profiling(True)
    a = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=30, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    prof = pcreg.last_profile()
    pcreg.set_profiling(False)
    b = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=30, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k], b[k]), k
    assert prof["voxel_map"] == 1 and prof["certified_queries"] > 0.95 * prof["nn_queries"], prof
    m.destroy()


@pytest.mark.parametrize("mode", ["plain-reject", "knn", "knn-reject", "weighted"])
def test_fused_kernel_equals_per_pass_kernels_and_brute_force(pcreg, monkeypatch, mode):
    """icp_fused.cu runs all passes of a hypothesis in one block; it must return the bits of the per-pass kernels
    (PCREG_FUSED=0: k_nn_vox + k_icp_update) and of the brute-force path, for every selection mode."""
    model = synth.make_model(80_000, 191)
    src, T_gt, c = synth.make_source(model, 1500, 0.3, 192)
    g = synth.rng(7)
    src = np.vstack([src, src[:200] + g.normal(0, 2.5, (200, 3))])            # some gross outliers
    T0 = synth.pose_grid(T_gt, c, 3, (2, 2, 2), 12.0, 2.0, 19)[:21]
    kw = {"plain-reject": dict(mode=pcreg.ICP_PLAIN, thDist2=4.0), "knn": dict(mode=pcreg.ICP_KNN),
          "knn-reject": dict(mode=pcreg.ICP_KNN, thDist2=9.0, k_frac=0.7),
          "weighted": dict(mode=pcreg.ICP_WEIGHTED, R_w=3.5, w_src=g.uniform(0.5, 1.0, src.shape[0]))}[mode]
    m = pcreg.Model(model, grid=True)
    monkeypatch.setenv("PCREG_FUSED", "1")                 # (by default only batches of >= one hypothesis per SM run fused)
    pcreg.set_profiling(True)
    a = pcreg.icp_batch(m, src, T0, iters=14, nn=pcreg.NN_GRID, return_idx=True, return_hist=True, **kw)
    prof = pcreg.last_profile()
    pcreg.set_profiling(False)
    assert prof["fused"] == 1 and prof["voxel_map"] == 1, prof
    monkeypatch.setenv("PCREG_FUSED", "0")
    b = pcreg.icp_batch(m, src, T0, iters=14, nn=pcreg.NN_GRID, return_idx=True, return_hist=True, **kw)
    monkeypatch.delenv("PCREG_FUSED")
    r = pcreg.icp_batch(m, src, T0, iters=14, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True, **kw)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k], b[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], b[k]), ("fused vs per-pass", k)
        assert np.array_equal(a[k], r[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], r[k]), ("fused vs brute", k)
    assert a["best"] == b["best"] == r["best"]
    m.destroy()


def test_fused_kernel_walks_queries_without_a_list(pcreg, monkeypatch):
    """Dropped lists (tiny cap) and queries outside the unpadded box: the fused kernel walks the pyramid for them in place."""
    model = synth.make_model(60_000, 23)
    src, T_gt, c = synth.make_source(model, 1200, 0.3, 24)
    T0 = synth.pose_grid(T_gt, c, 2, (2, 2, 2), 10.0, 3.0, 3)
    monkeypatch.setenv("PCREG_VOX_CAP", "8")
    m = pcreg.Model(model, grid=True, voxel_map=1, voxel_margin=-1.0)
    monkeypatch.delenv("PCREG_VOX_CAP")
    monkeypatch.setenv("PCREG_FUSED", "1")
    pcreg.set_profiling(True)
    a = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=10, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    prof = pcreg.last_profile()
    pcreg.set_profiling(False)
    assert prof["fused"] == 1 and prof["walked_queries"] > 0, prof
    r = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=10, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k], r[k]), k
    m.destroy()


def test_fused_kernel_frozen_and_short_sources(pcreg, monkeypatch):
    """Fewer than 3 usable correspondences freezes the pose (status 1); tiny sources (ns < block size, ns = 1)."""
    monkeypatch.setenv("PCREG_FUSED", "1")
    model = synth.make_model(20_000, 300)
    for ns in (1, 2, 37, 600):
        src, T_gt, c = synth.make_source(model, max(ns, 4), 0.3, 301)
        src = src[:ns]
        T0 = synth.pose_grid(T_gt, c, 2, (2, 1, 1), 5.0, 1.0, 3)
        m = pcreg.Model(model, grid=True)
        for kw in (dict(mode=pcreg.ICP_PLAIN, thDist2=1e-12), dict(mode=pcreg.ICP_KNN), dict(mode=pcreg.ICP_PLAIN)):
            a = pcreg.icp_batch(m, src, T0, iters=5, nn=pcreg.NN_GRID, return_idx=True, return_hist=True, **kw)
            r = pcreg.icp_batch(m, src, T0, iters=5, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True, **kw)
            for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
                assert np.array_equal(a[k], r[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], r[k]), (ns, kw, k)
        m.destroy()


def test_fused_kernel_is_the_default_for_batches_that_fill_the_gpu(pcreg):
    """>= one hypothesis per SM: the fused kernel runs by itself (no environment switch) and equals brute force."""
    model = synth.make_model(40_000, 61)
    src, T_gt, c = synth.make_source(model, 600, 0.3, 62)
    T0 = synth.pose_grid(T_gt, c, 4, (4, 4, 3), 10.0, 1.5, 3)              # 192 hypotheses
    m = pcreg.Model(model, grid=True)
    pcreg.set_profiling(True)
    a = pcreg.icp_batch(m, src, T0, mode=pcreg.ICP_KNN, iters=8, nn=pcreg.NN_GRID, return_idx=True, return_hist=True)
    prof = pcreg.last_profile()
    pcreg.set_profiling(False)
    assert prof["fused"] == 1, prof
    r = pcreg.icp_batch(m, src, T0[:24], mode=pcreg.ICP_KNN, iters=8, nn=pcreg.NN_BRUTE, return_idx=True, return_hist=True)
    for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
        assert np.array_equal(a[k][:24], r[k]), k
    m.destroy()


def test_band_limited_map_of_dense_models(pcreg, monkeypatch):
    """Dense models get lists only for the voxels within a band around the model (nn_vox.cu); queries beyond the band are
    walked.  Forced here on a small model: results stay exact on both sides of the band."""
    model = synth.make_model(60_000, 71)
    g = synth.rng(72)
    near = np.asarray(model[:3000], dtype=np.float64) + g.normal(0, 0.4, (3000, 3))
    far = np.asarray(model[:1500], dtype=np.float64) + g.normal(0, 6.0, (1500, 3))
    monkeypatch.setenv("PCREG_VOX_BAND", "1.5")
    m = pcreg.Model(model, grid=True, voxel_map=1)
    monkeypatch.delenv("PCREG_VOX_BAND")
    info = m.voxel_info()
    assert info["outside_band"] > 0 and abs(info["band"] - 1.5) < 1e-6 and info["listed"] > 0, info
    prof = _nn_equal(pcreg, m, model, np.vstack([near, far]))
    assert prof["certified_queries"] > 2000 and prof["walked_queries"] > 500, prof
    m.destroy()
