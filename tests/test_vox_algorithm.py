"""CPU check of the Voronoi voxel map ALGORITHM (tools/vox_sim.py restates the build and the query of
pcreg_b200/csrc/nn_vox.cu in numpy, FP32 where the kernels use FP32): the list of a voxel must contain the exact nearest
neighbour -- every exact tie included -- of every location inside the voxel, and the FP32 scan + FP64 decision must return
the FP64 brute-force answer (ties -> smallest index).  The CUDA kernels themselves are checked against the oracle in
tests/test_gpu_voxel_map.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import vox_sim  # noqa: E402
from pcreg_b200 import synth  # noqa: E402


def _brute(pts, x):
    d2 = ((pts - x) ** 2).sum(1)
    b = d2.min()
    return int(np.nonzero(d2 == b)[0].min()), float(b)


def _run(pts, s, lo, dims, queries):
    vox = vox_sim.build(pts, s, lo, dims)
    answered = 0
    for x in queries:
        r = vox_sim.query(pts, vox, s, lo, dims, x)
        if r is None:
            continue
        answered += 1
        assert r == _brute(pts, x), (x, r, _brute(pts, x))
    return vox, answered


def test_surface_patch_random_queries():
    g = np.random.default_rng(1)
    model = synth.make_model(60_000, 5).astype(np.float64)
    patch = model[np.abs(model - model[77]).max(1) < 5.0]
    s = 0.9
    lo = patch.min(0) - 1.0
    dims = tuple(int(np.ceil((patch.max(0) + 1.0 - lo)[k] / s)) for k in range(3))
    q = lo + g.uniform(0, 1, (4000, 3)) * np.array(dims) * s
    vox, answered = _run(patch, s, lo, dims, q)
    assert answered == len(q)
    lens = np.array([len(v) for v in vox.values() if v is not None])
    assert lens.size == len(vox) and lens.mean() < 20


def test_lattice_with_duplicates_exact_ties():
    ax = np.arange(6, dtype=np.float64)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    pts = np.column_stack([X.ravel(), Y.ravel(), Z.ravel()])
    g = np.random.default_rng(5)
    model = np.vstack([pts, pts])[g.permutation(2 * pts.shape[0])]
    s = 0.5
    lo = np.array([-1.0, -1.0, -1.0])
    dims = (14, 14, 14)
    q = np.vstack([pts[:80] + 0.5, pts[:40], pts[:40] + np.array([0.5, 0.0, 0.0]), lo + g.uniform(0, 7, (300, 3))])
    _, answered = _run(model, s, lo, dims, q)
    assert answered == len(q)


def test_dropped_lists_are_reported_not_wrong():
    """Points on a sphere, voxels around its centre: thousands of near-equidistant candidates -> lists over the cap are
    dropped (None) and such queries fall back to the walk; nothing is ever answered from a truncated list."""
    g = np.random.default_rng(3)
    d = g.standard_normal((3000, 3))
    pts = 10.0 * d / np.linalg.norm(d, axis=1, keepdims=True)
    s = 1.0
    lo = np.array([-2.0, -2.0, -2.0])
    dims = (4, 4, 4)
    vox = vox_sim.build(pts, s, lo, dims, base_cap=16)
    assert any(v is None for v in vox.values())
    for x in lo + g.uniform(0, 4, (300, 3)):
        r = vox_sim.query(pts, vox, s, lo, dims, x)
        if r is not None:
            assert r == _brute(pts, x)
