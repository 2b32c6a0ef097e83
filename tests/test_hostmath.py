"""CPU tests of the FP64 3x3 linear algebra the kernels inline (pcreg_b200/csrc/pcreg_math.cuh),
compiled host-only behind tests/hostmath.  Checked against numpy / the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pcreg_b200.build import build_hostmath  # noqa: E402

dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hm():
    lib = C.CDLL(build_hostmath())
    lib.hm_svd3.argtypes = [dp, dp, dp, dp]
    lib.hm_eigsym3.argtypes = [dp, dp, dp, C.c_int]
    lib.hm_kabsch_from_sums.argtypes = [dp, dp, dp, C.c_int, dp]
    lib.hm_mul4.argtypes = [dp, dp, dp]
    lib.hm_spacing.argtypes = [C.c_double]
    lib.hm_spacing.restype = C.c_double
    lib.hm_rank_from_sv.argtypes = [dp, C.c_longlong]
    lib.hm_rank_from_sv.restype = C.c_int
    lib.hm_det3.argtypes = [dp]
    lib.hm_det3.restype = C.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(dp)


def svd3(hm, A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    U, S, V = np.empty((3, 3)), np.empty(3), np.empty((3, 3))
    hm.hm_svd3(_p(A), _p(U), _p(S), _p(V))
    return U, S, V


def test_svd3_random_and_graded(hm):
    g = np.random.default_rng(0)
    for k in range(300):
        A = g.normal(size=(3, 3)) * 10.0 ** g.uniform(-6, 6)
        if k % 3 == 0:          # graded singular values
            Q1, _ = np.linalg.qr(g.normal(size=(3, 3)))
            Q2, _ = np.linalg.qr(g.normal(size=(3, 3)))
            A = Q1 @ np.diag([1.0, 10.0 ** -g.uniform(0, 10), 10.0 ** -g.uniform(0, 14)]) @ Q2.T
        U, S, V = svd3(hm, A)
        s_ref = np.linalg.svd(A, compute_uv=False)
        assert np.all(np.diff(S) <= 0)
        assert np.max(np.abs(S - s_ref)) <= 1e-14 * s_ref[0]
        assert np.max(np.abs(U @ np.diag(S) @ V.T - A)) <= 1e-14 * s_ref[0]
        assert np.max(np.abs(U.T @ U - np.eye(3))) < 1e-13 and np.max(np.abs(V.T @ V - np.eye(3))) < 1e-13


def test_svd3_rank_deficient(hm):
    g = np.random.default_rng(1)
    u, v = g.normal(size=3), g.normal(size=3)
    for A in (np.outer(u, v), np.zeros((3, 3)), np.outer(u, v) + np.outer(g.normal(size=3), g.normal(size=3))):
        U, S, V = svd3(hm, A)
        assert np.max(np.abs(U @ np.diag(S) @ V.T - A)) <= 1e-14 * max(S[0], 1e-300)
        assert np.max(np.abs(U.T @ U - np.eye(3))) < 1e-13


def test_polar_factor_matches_lapack(hm):
    """R = V*U' (estimateTransform.m:62) is what parity depends on: compare with numpy's SVD."""
    g = np.random.default_rng(2)
    for _ in range(200):
        H = g.normal(size=(3, 3)) * 100
        U, S, V = svd3(hm, H)
        Un, _, Vtn = np.linalg.svd(H)
        assert np.max(np.abs(V @ U.T - Vtn.T @ Un.T)) < 1e-12


def test_eigsym3_matches_eigh(hm):
    g = np.random.default_rng(3)
    for _ in range(300):
        B = g.normal(size=(3, 3))
        A = np.ascontiguousarray(B @ B.T * 10.0 ** g.uniform(-3, 3))
        for d in (+1, -1):
            w, V = np.empty(3), np.empty((3, 3))
            hm.hm_eigsym3(_p(A), _p(w), _p(V), d)
            wr = np.linalg.eigvalsh(A)
            assert np.allclose(np.sort(w), wr, rtol=1e-12, atol=1e-14 * wr[-1])
            assert np.all(np.diff(w) * d >= 0)
            assert np.max(np.abs(A @ V - V * w)) <= 1e-13 * wr[-1]
            assert np.max(np.abs(V.T @ V - np.eye(3))) < 1e-13


def _sums(q, m, w, pq, pm):
    qq, mm = q - pq, m - pm
    s = np.zeros(17)
    s[0] = w.sum()
    s[1:4] = (w[:, None] * qq).sum(0)
    s[4:7] = (w[:, None] * mm).sum(0)
    s[7:16] = ((w[:, None] * qq).T @ mm).reshape(9)
    return s


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(4, 200), st.booleans())
def test_kabsch_from_sums_matches_oracle_estimateTransform(seed, n, weighted):
    lib = C.CDLL(build_hostmath())
    lib.hm_kabsch_from_sums.argtypes = [dp, dp, dp, C.c_int, dp]
    g = np.random.default_rng(seed)
    q = g.normal(0, 8, (n, 3)) + g.uniform(-60, 60, 3)
    Rt = np.linalg.qr(g.normal(size=(3, 3)))[0]
    m = q @ Rt + g.uniform(-20, 20, 3) + g.normal(0, 0.2, (n, 3))
    w = g.uniform(0.1, 2.0, n) if weighted else np.ones(n)
    pq, pm = q[0].copy(), m[0].copy()
    s = _sums(q, m, w, pq, pm)
    dT = np.empty(16)
    lib.hm_kabsch_from_sums(_p(s), _p(pq), _p(pm), 0, _p(dT))
    dT = dT.reshape(4, 4)
    if weighted:
        from oracle.icp import _weighted_kabsch
        ref = _weighted_kabsch(m, q, w, False)
    else:
        ref = oracle.estimateTransform(m, q, rank_guard=False)
    assert np.linalg.norm(dT[:3, :3] - ref[:3, :3]) < 1e-10
    assert np.linalg.norm(dT[3, :3] - ref[3, :3]) < 1e-9 * max(1.0, np.linalg.norm(ref[3, :3]))
    assert np.allclose(dT[:, 3], [0, 0, 0, 1])


def test_spacing_and_rank(hm):
    for x in (1.0, 1.5, 2.0, 1e-300, 3.7e10, 2.0 ** 52, 0.1):
        assert hm.hm_spacing(x) == np.spacing(x)
    s = np.array([5.0, 1e-3, 1e-15])
    assert hm.hm_rank_from_sv(_p(s), 3) == 2
    assert hm.hm_rank_from_sv(_p(s), 10 ** 6) == 2
    s = np.array([5.0, 1e-3, 1e-14])
    assert hm.hm_rank_from_sv(_p(s), 3) == 3


def test_mul4_and_det3(hm):
    g = np.random.default_rng(5)
    A, B = g.normal(size=(4, 4)), g.normal(size=(4, 4))
    Cm = np.empty((4, 4))
    hm.hm_mul4(_p(A), _p(B), _p(Cm))
    assert np.allclose(Cm, A @ B, rtol=1e-14, atol=1e-14)
    M = np.ascontiguousarray(g.normal(size=(3, 3)))
    assert abs(hm.hm_det3(_p(M)) - np.linalg.det(M)) < 1e-13
