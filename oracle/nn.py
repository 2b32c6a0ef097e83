"""Oracle nearest-neighbour search (FP64, K=1, ties -> smallest index).  TEST INFRASTRUCTURE ONLY.

The reference has no NN correspondence code; semantics follow MATLAB's documented knnsearch
 -- PARITY UNPINNED.  Two implementations that must agree:
  nn_brute  : the definition -- exhaustive scan in C (oracle/nn_brute.c), numpy fallback.
  nn_kdtree : scipy cKDTree candidates made exact with the same d2 formula; this is what the
              CPU baseline times (MATLAB's knnsearch would also use a kd-tree for 3-D data).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/nn_brute.c -> oracle/_build/liboracle_nn.so (gcc + OpenMP)."""
    out = os.path.join(_HERE, "_build", "liboracle_nn.so")
    src = os.path.join(_HERE, "nn_brute.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_nn.so")
        if not os.path.exists(path):
            try:
                build_c_oracle()
            except Exception:
                return None
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.oracle_nn_brute_f64.argtypes = [dp, dp, dp, ctypes.c_int64, dp, dp, dp, ctypes.c_int64, ip, dp]
        lib.oracle_nn_brute_f64.restype = None
        lib.oracle_nn_second_f64.argtypes = [dp, dp, dp, ctypes.c_int64, dp, dp, dp, ctypes.c_int64, ip, dp]
        lib.oracle_nn_second_f64.restype = None
        _LIB = lib
    return _LIB


def d2_exact(m, q):
    """((mx-qx)^2 + (my-qy)^2) + (mz-qz)^2, the one distance formula used everywhere."""
    dx = m[..., 0] - q[..., 0]
    dy = m[..., 1] - q[..., 1]
    dz = m[..., 2] - q[..., 2]
    return (dx * dx + dy * dy) + dz * dz


def _soa(a):
    a = np.asarray(a, dtype=np.float64)
    return [np.ascontiguousarray(a[:, c]) for c in range(3)]


def nn_brute(model, q, use_c: bool = True):
    """Exhaustive FP64 1-NN: returns (idx int32 [nq], d2 float64 [nq])."""
    model = np.asarray(model, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    nq = q.shape[0]
    lib = _lib() if use_c else None
    if lib is not None:
        m3, q3 = _soa(model), _soa(q)
        idx = np.empty(nq, dtype=np.int32)
        d2 = np.empty(nq, dtype=np.float64)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.oracle_nn_brute_f64(*[a.ctypes.data_as(dp) for a in m3], model.shape[0],
                                *[a.ctypes.data_as(dp) for a in q3], nq,
                                idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), d2.ctypes.data_as(dp))
        return idx, d2
    idx = np.empty(nq, dtype=np.int32)
    d2 = np.empty(nq, dtype=np.float64)
    blk = max(1, int(4e6 // max(model.shape[0], 1)))
    for s in range(0, nq, blk):
        D = d2_exact(model[None, :, :], q[s:s + blk, None, :])
        j = np.argmin(D, axis=1)          # first minimum = smallest index
        idx[s:s + blk] = j
        d2[s:s + blk] = D[np.arange(j.shape[0]), j]
    return idx, d2


def nn_second(model, q, idx):
    """Second-best squared distance (best index excluded) -- tells tests where index parity is
    required (gap > 1e-9 relative, BASELINE.json north_star)."""
    model = np.asarray(model, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    lib = _lib()
    nq = q.shape[0]
    out = np.empty(nq, dtype=np.float64)
    if lib is not None:
        m3, q3 = _soa(model), _soa(q)
        dp = ctypes.POINTER(ctypes.c_double)
        skip = np.ascontiguousarray(idx, dtype=np.int32)
        lib.oracle_nn_second_f64(*[a.ctypes.data_as(dp) for a in m3], model.shape[0],
                                 *[a.ctypes.data_as(dp) for a in q3], nq,
                                 skip.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), out.ctypes.data_as(dp))
        return out
    for i in range(nq):
        D = d2_exact(model, q[i][None, :])
        D[idx[i]] = np.inf
        out[i] = D.min()
    return out


class KDTreeNN:
    """cKDTree wrapper whose answers are made exact (same d2 formula, ties -> smallest index)."""

    def __init__(self, model):
        from scipy.spatial import cKDTree
        self.model = np.ascontiguousarray(model, dtype=np.float64)
        self.tree = cKDTree(self.model)

    def query(self, q, workers: int = -1):
        q = np.ascontiguousarray(q, dtype=np.float64)
        d, j = self.tree.query(q, k=2, workers=workers)
        j0 = j[:, 0].astype(np.int64)
        d2 = d2_exact(self.model[j0], q)
        # near ties (or kd-tree rounding): re-decide among everything within the band, exactly
        near = (d[:, 1] - d[:, 0]) <= 1e-9 * np.maximum(d[:, 1], 1e-300)
        for i in np.nonzero(near)[0]:
            cand = self.tree.query_ball_point(q[i], r=d[i, 1] * (1 + 1e-9) + 1e-300)
            cand = np.asarray(sorted(cand), dtype=np.int64)
            dd = d2_exact(self.model[cand], q[i][None, :])
            k = int(np.argmin(dd))       # first minimum over index-sorted candidates
            j0[i] = cand[k]
            d2[i] = dd[k]
        return j0.astype(np.int32), d2


def nn_kdtree(model, q, workers: int = -1):
    return KDTreeNN(model).query(q, workers=workers)
