"""Seeded synthetic inputs for tests and bench.py.  numpy only; this is input
generation, not part of the compute path.

Sizing constants come from the reference: voxel spacing 0.1953 x 0.1953 x 0.34 mm
(createExampleCrops.m:64-66), model crop about 101.6 x 55.9 x 98.9 mm (GetPointcloudFromModel.m:47-53),
surface patch half-diagonal about 10.6 mm (debugCompleteExperiment.m:8), test translation [13,25,-17]
(slideMatchingWindow.m:35-36), 2 mm pose-grid pitch (completeExperiment.m:79).
"""
from __future__ import annotations

import numpy as np

VOXEL = np.array([0.1953, 0.1953, 0.34])
BOX = np.array([101.6, 55.9, 98.9])


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def rot_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def rot_xyz(e):
    """R = Rx(e1) Ry(e2) Rz(e3) (eul2rotm 'XYZ' as used by pcRigidBodyTF.m:13)."""
    return rot_axis_angle([1, 0, 0], e[0]) @ rot_axis_angle([0, 1, 0], e[1]) @ rot_axis_angle([0, 0, 1], e[2])


def make_T(R, t):
    """Row-vector 4x4: [p 1] @ T = p @ R + t."""
    T = np.eye(4)
    T[:3, :3] = R
    T[3, :3] = t
    return T


def apply_T(pts, T):
    pts = np.asarray(pts, dtype=np.float64)
    return pts @ T[:3, :3] + T[3, :3]


def invert_T(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[3, :3] = -T[3, :3] @ T[:3, :3].T
    return Ti


def make_model(n, seed, dtype=np.float32):
    """CT-surface-like closed star-shaped surface, voxel-lattice snapped + small jitter, class single."""
    g = rng(seed)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    theta = np.arccos(np.clip(d[:, 2], -1, 1))
    phi = np.arctan2(d[:, 1], d[:, 0])
    a = g.standard_normal(6)
    ph = g.uniform(0, 2 * np.pi, 12)
    r = 40.0 + np.zeros(n)
    for k in range(6):
        r += 3.0 * a[k] / (1 + 0.5 * k) * np.sin((k + 1) * theta + ph[k]) * np.cos((k % 3 + 1) * phi + ph[6 + k])
    p = d * r[:, None]
    ext = p.max(axis=0) - p.min(axis=0)
    p = (p - p.min(axis=0)) * (BOX / ext)              # fill the reference's crop box
    p = np.round(p / VOXEL) * VOXEL                    # voxel lattice
    p += g.uniform(-0.05, 0.05, p.shape)               # de-duplicating jitter
    return np.ascontiguousarray(p.astype(dtype))


def make_source(model, ns, sigma, seed, min_radius=10.6):
    """Sparse 'stereo' patch: ns noisy model points around a random surface point, moved by the inverse of
    a ground-truth pose.  Returns (src float64 [ns,3], T_gt, patch_centre_in_model_frame)."""
    g = rng(seed)
    m = np.asarray(model, dtype=np.float64)
    c = m[g.integers(0, m.shape[0])]
    d2 = ((m - c) ** 2).sum(axis=1)
    need = min(m.shape[0], 4 * ns)
    rad2 = max(np.partition(d2, need - 1)[need - 1], min_radius ** 2)
    cand = np.nonzero(d2 <= rad2)[0]
    pick = g.choice(cand, size=ns, replace=cand.size < ns)
    patch = m[pick] + g.normal(0.0, sigma, (ns, 3))
    T_gt = make_T(rot_xyz(g.uniform(0, 2 * np.pi, 3)), np.array([13.0, 25.0, -17.0]))
    src = apply_T(patch, invert_T(T_gt))
    return np.ascontiguousarray(src), T_gt, c


def perturb_pose(T_gt, centre, R_p, t_p):
    """T0 = T_gt followed by a rotation R_p about `centre` (model frame) and a shift t_p."""
    P = make_T(R_p, centre - centre @ R_p + t_p)
    return T_gt @ P


def pose_grid(T_gt, centre, n_rot, n_trans_xyz, max_deg, pitch, seed):
    """slideMatchingWindow / completeExperiment style multi-start grid: n_rot random rotations (angle <=
    max_deg about random axes through the patch centre) x a translation lattice (nx, ny, nz) of `pitch` mm
    (a scalar, or one pitch per axis)."""
    g = rng(seed)
    rots = [np.eye(3)]
    for _ in range(n_rot - 1):
        rots.append(rot_axis_angle(g.standard_normal(3), np.deg2rad(g.uniform(0, max_deg))))
    nx, ny, nz = n_trans_xyz
    px, py, pz = np.broadcast_to(np.asarray(pitch, dtype=np.float64), (3,))
    gx = (np.arange(nx) - (nx - 1) / 2) * px
    gy = (np.arange(ny) - (ny - 1) / 2) * py
    gz = (np.arange(nz) - (nz - 1) / 2) * pz
    out = []
    for R in rots:
        for x in gx:
            for y in gy:
                for z in gz:
                    out.append(perturb_pose(T_gt, centre, R, np.array([x, y, z])))
    return np.stack(out)


def make_ransac_problem(P, inlier_frac, sigma, seed):
    """Putative matches: pts1 = model-side keypoints, pts2 = surface-side; a fraction are true
    correspondences under a random rigid T (plus noise), the rest uniform in the box."""
    g = rng(seed)
    p2 = g.uniform(0, 1, (P, 3)) * BOX
    T = make_T(rot_xyz(g.uniform(0, 2 * np.pi, 3)), np.array([13.0, 25.0, -17.0]))
    p1 = apply_T(p2, T) + g.normal(0, sigma, (P, 3))
    n_out = P - int(round(inlier_frac * P))
    out_idx = g.choice(P, size=n_out, replace=False)
    p1[out_idx] = g.uniform(0, 1, (n_out, 3)) * BOX
    return np.ascontiguousarray(p1), np.ascontiguousarray(p2), T


def make_triplets(P, n, seed):
    """n sample triplets of distinct indices (the host-side stand-in for randperm(P)(1:3), ransac.m:42-43)."""
    g = rng(seed)
    t = np.empty((n, 3), dtype=np.int32)
    t[:, 0] = g.integers(0, P, n)
    t[:, 1] = (t[:, 0] + 1 + g.integers(0, P - 1, n)) % P
    third = g.integers(0, P - 2, n)
    lo = np.minimum(t[:, 0], t[:, 1])
    hi = np.maximum(t[:, 0], t[:, 1])
    third = third + (third >= lo)
    third = third + (third >= hi)
    t[:, 2] = third
    return t


def make_neighbourhoods(nb, seed, nmin=500, nmax=6000, radius=3.5, dtype=np.float64):
    """Surface-like neighbourhoods: points of a noisy curved sheet inside a ball of `radius`
    (getLocalPoints output: 500..6000 points, GetSphericalDescriptors.m:133-139)."""
    g = rng(seed)
    out = []
    for _ in range(nb):
        n = int(g.integers(nmin, nmax + 1))
        uv = g.uniform(-1, 1, (3 * n, 2)) * radius
        k1, k2 = g.uniform(-0.15, 0.15, 2)
        z = k1 * uv[:, 0] ** 2 + k2 * uv[:, 1] ** 2 + 0.2 * uv[:, 0] + g.normal(0, 0.05, 3 * n)
        p = np.column_stack([uv, z])
        p = p[np.linalg.norm(p, axis=1) < radius][:n]
        R = rot_xyz(g.uniform(0, 2 * np.pi, 3))
        out.append(np.ascontiguousarray((p @ R + g.uniform(-30, 30, 3)).astype(dtype)))
    return out
