"""world_size-2 gloo test of the N>1 path's host logic: hypothesis sharding, the all-gather of result
records and the first-index arg-min (pcreg_b200/sharded.py).  The per-rank compute is a stand-in here
(the oracle on a tiny problem) -- the CUDA call is exercised by the -m gpu tests."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from pcreg_b200 import sharded, synth
    model = synth.make_model(3000, 5)
    src, T_gt, c = synth.make_source(model, 120, 0.2, 6)
    T0 = synth.pose_grid(T_gt, c, 1, (3, 1, 1), 0.0, 1.0, 7)
    T0 = np.concatenate([T0, T0[:2]])            # 5 hypotheses (odd: ragged shards), two exact duplicates -> rmse ties
    nn = oracle.nn.KDTreeNN(np.asarray(model, dtype=np.float64))

    def local_fn(T0_local):
        res = [oracle.icp_single(model, src, T, mode=oracle.ICP_KNN, iters=4, nn=nn) for T in T0_local]
        return dict(T=np.stack([r["T"] for r in res]), rmse=np.array([r["rmse"] for r in res]),
                    n_used=np.array([r["n_used"] for r in res]), status=np.array([r["status"] for r in res]))

    out = sharded.icp_batch_sharded(local_fn, T0)
    ref = oracle.icp_batch(model, src, T0, mode=oracle.ICP_KNN, iters=4)
    ok = (np.array_equal(out["T"], ref["T"]) and np.array_equal(out["rmse"], ref["rmse"])
          and np.array_equal(out["n_used"], ref["n_used"]) and out["best"] == ref["best"])
    lo, hi, per = sharded.shard_range(5, rank, world)
    q.put((rank, ok, out["best"], (lo, hi, per)))
    dist.destroy_process_group()


def test_sharded_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] and got[1][1]
    assert got[0][2] == got[1][2]                # every rank agrees on the winner
    assert got[0][3] == (0, 3, 3) and got[1][3] == (3, 5, 3)


def test_first_argmin_rules():
    from pcreg_b200.sharded import first_argmin, shard_range
    nan = float("nan")
    assert first_argmin(torch.tensor([nan, 2.0, 1.0, 1.0, nan], dtype=torch.float64)) == 2
    assert first_argmin(torch.tensor([nan, nan], dtype=torch.float64)) == -1
    assert first_argmin(torch.tensor([0.5], dtype=torch.float64)) == 0
    assert [shard_range(10, r, 4)[:2] for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [shard_range(2, r, 4)[:2] for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
