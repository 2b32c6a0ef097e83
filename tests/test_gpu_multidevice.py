"""Multi-GPU behind the C ABI (include/pcreg.h: pcreg_init(devices, ndev > 1)): the model is replicated on every selected
device, pcreg_icp_batch shards its hypotheses and pcreg_ransac_batch its windows over the devices inside the library
(one host thread per device) -- the replacement of the reference's parfor (slideMatchingWindow_v2.m:178,
completeExperiment.m:265).  Results must be bit-identical to the single-device call for any device count.

On a box with one GPU the same code path (worker threads, per-slot pools / streams / replicas) runs with several slots on
that one device (PCREG_ALLOW_DUP_DEVICES=1); with >= 2 GPUs it runs on distinct devices."""
import os

import numpy as np
import pytest

from pcreg_b200 import synth

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]                              # slots sharing device 0 (any box)
    if n >= 2:
        sets.append(list(range(min(n, 8))))                 # distinct devices
    return sets


@pytest.fixture()
def multi(pcreg):
    os.environ["PCREG_ALLOW_DUP_DEVICES"] = "1"
    yield pcreg
    pcreg.init(0)                                           # back to the session's single-device state


def test_icp_batch_sharded_over_devices_is_bit_identical(multi):
    P = multi
    model = synth.make_model(60_000, 91)
    src, T_gt, c = synth.make_source(model, 1100, 0.3, 92)
    T0 = synth.pose_grid(T_gt, c, 4, (2, 2, 2), 8.0, 1.5, 9)[:29]           # odd count: uneven shares
    g = synth.rng(5)
    cases = [dict(mode=P.ICP_KNN, nn=P.NN_GRID), dict(mode=P.ICP_WEIGHTED, nn=P.NN_GRID, w_src=g.uniform(0.5, 1, src.shape[0])),
             dict(mode=P.ICP_PLAIN, thDist2=4.0, nn=P.NN_BRUTE)]
    P.init(0)
    m = P.Model(model, grid=True)
    ref = [P.icp_batch(m, src, T0, iters=10, return_idx=True, return_hist=True, **kw) for kw in cases]
    m.destroy()
    for devs in _device_sets():
        P.init(devs)
        assert P.device_count() == len(devs)
        m = P.Model(model, grid=True)
        for kw, r in zip(cases, ref):
            a = P.icp_batch(m, src, T0, iters=10, return_idx=True, return_hist=True, **kw)
            for k in ("T", "idx", "rmse", "rmse_hist", "n_used", "status"):
                assert np.array_equal(a[k], r[k]), (devs, kw, k)
            assert a["best"] == r["best"]
        # fewer hypotheses than devices
        a = P.icp_batch(m, src, T0[:1], iters=5, mode=P.ICP_KNN, nn=P.NN_GRID)
        assert np.isfinite(a["rmse"]).all() and a["best"] == 0
        m.destroy()


def test_ransac_batch_sharded_over_devices_is_bit_identical(multi):
    P = multi
    g = synth.rng(12)
    wins = []
    for w in range(7):
        p1, p2, _ = synth.make_ransac_problem(int(g.integers(40, 300)), 0.3, 0.15, 100 + w)
        wins.append((p1, p2))
    wins.insert(3, (np.zeros((2, 3)), np.zeros((2, 3))))                      # a window with fewer than 3 pairs -> []
    coef = dict(thDist=0.3, thInlrRatio=0.08, REFINE=True, iterNum=2000)
    seeds = [977 * (w + 1) for w in range(len(wins))]
    p1s, p2s = [w[0] for w in wins], [w[1] for w in wins]
    P.init(0)
    ref = P.ransac_batch(p1s, p2s, coef, seeds=seeds)
    assert ref[3]["T"] is None and any(r["T"] is not None for r in ref)
    for devs in _device_sets():
        P.init(devs)
        out = P.ransac_batch(p1s, p2s, coef, seeds=seeds)
        for a, r in zip(out, ref):
            assert (a["T"] is None) == (r["T"] is None)
            for k in ("inlierIdx", "numSuccess", "maxInliers", "best"):
                assert np.array_equal(np.asarray(a[k]), np.asarray(r[k])), (devs, k)
            if a["T"] is not None:
                assert np.array_equal(a["T"], r["T"]), devs
