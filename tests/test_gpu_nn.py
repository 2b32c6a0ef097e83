"""Parity of the CUDA nearest-neighbour paths (brute force and grid) against the FP64 oracle.
Bar: indices and squared distances BIT-EXACT (integer / index work; the oracle formula is reproduced
operation by operation), including exact ties -> smallest index."""
import numpy as np
import pytest

import oracle
from pcreg_b200 import synth

pytestmark = pytest.mark.gpu


def _check(pcreg, model, q, grid_kw=None):
    """brute force, the grid kernels (voxel_map=-1) and the Voronoi voxel map (forced: voxel_map=1) against the oracle"""
    oi, od = oracle.nn_brute(model, q)
    for vm in (-1, 1):
        m = pcreg.Model(model, grid=True, voxel_map=vm, **(grid_kw or {}))
        for kind in ((pcreg.NN_BRUTE, pcreg.NN_GRID) if vm < 0 else (pcreg.NN_GRID,)):
            gi, gd = m.nn_search(q, kind)
            bad = np.nonzero(gi != oi)[0]
            assert bad.size == 0, "kind %d (voxel_map %d): %d index mismatches, first at %s: got %s want %s (d2 %s vs %s)" % (
                kind, vm, bad.size, bad[:5], gi[bad[:5]], oi[bad[:5]], gd[bad[:5]], od[bad[:5]])
            assert np.array_equal(gd, od), "kind %d (voxel_map %d): d2 not bit-exact (max rel %g)" % (kind, vm, np.max(np.abs(gd - od) / od))
        m.destroy()


def test_nn_surface_model_single(pcreg):
    model = synth.make_model(50_000, 1001)                   # class single, like upsampleMesh.m:21
    src, T_gt, _ = synth.make_source(model, 2000, 1.0, 11)
    q = synth.apply_T(src, T_gt)
    _check(pcreg, model, q)


def test_nn_far_queries_and_outside_bbox(pcreg):
    model = synth.make_model(30_000, 7)
    g = synth.rng(3)
    q = np.vstack([g.uniform(-80, 200, (1500, 3)), np.asarray(model[:100], dtype=np.float64),     # exact hits (d2 = 0)
                   np.asarray(model[:50], dtype=np.float64) + 1e-7])
    _check(pcreg, model, q)


def test_nn_exact_ties_smallest_index(pcreg):
    # lattice model with every point duplicated, queries at cell centres: 8-fold (x2) exact ties
    ax = np.arange(12, dtype=np.float64)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    pts = np.column_stack([X.ravel(), Y.ravel(), Z.ravel()])
    g = synth.rng(5)
    model = np.vstack([pts, pts])[g.permutation(2 * pts.shape[0])]
    q = np.vstack([pts[:600] + 0.5, pts[:300], pts[:300] + np.array([0.5, 0.0, 0.0])])
    _check(pcreg, model, q)


def test_nn_double_model_large_coordinates(pcreg):
    g = synth.rng(9)
    model = g.normal(0, 30, (20_000, 3)) + np.array([500.0, -300.0, 1000.0])      # class double, off-centre
    q = g.normal(0, 35, (3000, 3)) + np.array([500.0, -300.0, 1000.0])
    _check(pcreg, model, q)


def test_nn_tiny_models(pcreg):
    g = synth.rng(2)
    for n in (1, 2, 7, 1023, 1024, 1025):
        model = g.uniform(-1, 1, (n, 3))
        q = g.uniform(-2, 2, (257, 3))
        _check(pcreg, model, q)


def test_nn_degenerate_planar_model(pcreg):
    g = synth.rng(4)
    model = np.column_stack([g.uniform(0, 50, 5000), g.uniform(0, 50, 5000), np.zeros(5000)])
    q = np.column_stack([g.uniform(-5, 55, 1000), g.uniform(-5, 55, 1000), g.uniform(-3, 3, 1000)])
    _check(pcreg, model, q)


def test_nn_many_queries_uses_wide_kernel(pcreg):
    # > 148*256*8 queries exercises the 8-queries-per-thread instantiation
    model = synth.make_model(4096, 21)
    g = synth.rng(22)
    q = g.uniform(0, 100, (320_000, 3))
    oi, od = oracle.nn_brute(model, q)
    m = pcreg.Model(model, grid=False)
    gi, gd = m.nn_search(q, pcreg.NN_BRUTE)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    m.destroy()


def test_grid_requires_build(pcreg):
    m = pcreg.Model(synth.make_model(2000, 1), grid=False)
    with pytest.raises(pcreg.PcregError):
        m.nn_search(np.zeros((4, 3)), pcreg.NN_GRID)
    m.destroy()
