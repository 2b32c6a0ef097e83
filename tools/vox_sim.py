"""CPU replay of the Voronoi voxel map build and query of pcreg_b200/csrc/nn_vox.cu (numpy, FP32 where the kernels
use FP32).  Test / design infrastructure only -- never imported by the product.

It restates the algorithm (nine-pivot bisector test, top-down refinement, FP32 list scan with an error band and an FP64
decision) so that its invariants can be checked without a GPU (tests/test_vox_algorithm.py):
  * the list of a voxel contains the true nearest neighbour (and every exact tie) of every location inside it;
  * the FP32 scan + FP64 decision returns exactly the brute-force FP64 answer.

    python tools/vox_sim.py            # list-length statistics on a C3-like model patch
"""
from __future__ import annotations

import itertools

import numpy as np

f32 = np.float32
CORNERS = np.array([(0, 0, 0)] + list(itertools.product([1, -1], repeat=3)), dtype=np.float32)      # centre + 8 corners


def keep_mask(a, piv, e):
    """a [n,3] float32 (p - c), piv [9,3] float32: the nine-pivot bisector test of vox_keep()."""
    a = a.astype(f32)
    na = (a * a).sum(1, dtype=f32)
    keep = np.ones(a.shape[0], dtype=bool)
    tol_abs = f32(1e-6) * e * e
    for j in range(9):
        n0 = (piv[j] * piv[j]).sum(dtype=f32)
        l1 = np.abs(a - piv[j]).sum(1, dtype=f32)
        keep &= (na - n0) <= (f32(2) * e * l1 + f32(1e-5) * (na + n0) + tol_abs)
    return keep


def pivots(a, e):
    """arg-min of |a - r_j|^2 over the list for r_j = centre and the eight corners (pivot_update())."""
    out = np.empty((9, 3), dtype=f32)
    for j in range(9):
        d = ((a - CORNERS[j] * e) ** 2).sum(1)
        out[j] = a[int(np.argmin(d))]
    return out


def filter_list(pts, ids, centre, edge):
    """list of a voxel (centre, edge) filtered out of the candidate ids."""
    e = f32(0.5 * edge * (1.0 + 1e-4))
    a = (pts[ids] - centre).astype(f32)
    piv = pivots(a, e)
    return ids[keep_mask(a, piv, e)]


def build(pts, s, origin, dims, base_cap=64, top_dim=16):
    """Returns dict voxel (ix,iy,iz) -> ids (or None where the list was dropped) for the finest level."""
    pts = np.asarray(pts, dtype=np.float64)
    ld = [tuple(dims)]
    while max(ld[-1]) > top_dim:
        ld.append(tuple((d + 1) // 2 for d in ld[-1]))
    if len(ld) == 1:
        ld.append(tuple((d + 1) // 2 for d in ld[-1]))
    L = len(ld)
    allids = np.arange(pts.shape[0])
    parent = None
    for l in range(L - 1, -1, -1):
        cs = s * (1 << l)
        maxlen = min(pts.shape[0], base_cap * 4 ** l)
        cur = {}
        for ix in range(ld[l][0]):
            for iy in range(ld[l][1]):
                for iz in range(ld[l][2]):
                    centre = origin + (np.array([ix, iy, iz]) + 0.5) * cs
                    if parent is None:
                        src = allids
                    else:
                        src = parent[(ix // 2, iy // 2, iz // 2)]
                    if src is None or len(src) == 0:
                        cur[(ix, iy, iz)] = None
                        continue
                    lst = filter_list(pts, src, centre, cs)
                    cur[(ix, iy, iz)] = lst if len(lst) <= maxlen else None
        parent = cur
    return parent


def query(pts, vox, s, origin, dims, q):
    """k_nn_vox(): returns (idx, d2) or None when the voxel has no list / q is outside."""
    u = (q - origin) / s
    if not (np.all(u >= 0) and np.all(u < np.array(dims))):
        return None
    iv = tuple(int(v) for v in u)
    lst = vox[iv]
    if lst is None:
        return None
    c = origin + (np.array(iv) + 0.5) * s
    x = (q - c).astype(f32)
    a = (pts[lst] - c).astype(f32)
    dd = a - x
    d32 = (dd[:, 2] * dd[:, 2] + (dd[:, 1] * dd[:, 1] + dd[:, 0] * dd[:, 0])).astype(f32)
    m1 = d32.min()
    thr = f32(m1 * f32(3e-6) + m1) + f32(0.8e-6 * s * s)
    band = lst[d32 <= thr]
    d64 = ((pts[band] - q) ** 2).sum(1)
    best = d64.min()
    return int(band[d64 == best].min()), float(best)


if __name__ == "__main__":
    import sys
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    from pcreg_b200 import synth
    model = synth.make_model(1_000_000, 1003).astype(np.float64)
    c = model[1234]
    patch = model[np.abs(model - c).max(1) < 3.0]
    s = 0.2
    lo = patch.min(0) - 0.5
    dims = tuple(int(np.ceil((patch.max(0) + 0.5 - lo)[k] / s)) for k in range(3))
    print("patch", patch.shape[0], "points, voxels", dims)
    vox = build(patch, s, lo, dims)
    lens = np.array([len(v) for v in vox.values() if v is not None])
    print("listed %d of %d voxels, mean list %.1f, p50/p90/p99/max %s" % (lens.size, len(vox), lens.mean(), np.percentile(lens, [50, 90, 99, 100])))
