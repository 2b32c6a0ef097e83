"""Multi-GPU multi-start ICP: one process per GPU, hypotheses sharded, model replicated.

The reference runs independent hypotheses under parfor (slideMatchingWindow_v2.m:178,
completeExperiment.m:265, completeExperimentFast.m:201): there is no exchange during the iterations.
The only collective is the final all-gather of one fixed-size record per hypothesis
{rmse, T[16], n_used, status} followed by a first-index arg-min (ties -> smallest global hypothesis
index, mirroring MATLAB max/min and ransac.m:69-73).  torch.distributed is the plumbing: NCCL over
NVLink on GPUs, gloo in the CPU tests (where `local_fn` is a stand-in for the CUDA call).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

RECORD = 19   # rmse, T[16], n_used, status  (float64 each; the int fields are exactly representable)


def shard_range(nhyp: int, rank: int, world: int):
    """Contiguous ranges of ceil(nhyp/world) hypotheses (equal work: fixed iteration count)."""
    per = (nhyp + world - 1) // world
    lo = min(nhyp, rank * per)
    hi = min(nhyp, lo + per)
    return lo, hi, per


def pack_records(T, rmse, n_used, status, per: int, device) -> torch.Tensor:
    """[n_local] results -> [per, RECORD] float64 records, padded with NaN-rmse rows."""
    n = rmse.shape[0]
    rec = torch.full((per, RECORD), float("nan"), dtype=torch.float64, device=device)
    if n:
        rec[:n, 0] = torch.as_tensor(rmse, dtype=torch.float64, device=device)
        rec[:n, 1:17] = torch.as_tensor(T, dtype=torch.float64, device=device).reshape(n, 16)
        rec[:n, 17] = torch.as_tensor(n_used, device=device).to(torch.float64)
        rec[:n, 18] = torch.as_tensor(status, device=device).to(torch.float64)
    return rec


def first_argmin_t(rmse: torch.Tensor) -> torch.Tensor:
    """First index of the minimum as a 0-d int64 tensor ON THE DEVICE (no host synchronisation): NaN never wins,
    -1 if everything is NaN."""
    n = rmse.shape[0]
    v = torch.where(torch.isnan(rmse), torch.full_like(rmse, float("inf")), rmse)
    m = v.min()
    ar = torch.arange(n, dtype=torch.int64, device=rmse.device)
    cand = torch.where(v == m, ar, torch.full_like(ar, n))                 # ties -> smallest index
    best = cand.min()
    return torch.where(torch.isinf(m), torch.full_like(best, -1), best)


def first_argmin(rmse: torch.Tensor) -> int:
    """First index of the minimum, NaN never wins; -1 if everything is NaN (one host read)."""
    return int(first_argmin_t(rmse))


def gather_and_pick(rec_local: torch.Tensor, nhyp: int, per: int, group=None, on_device=False):
    """all_gather the per-rank records and pick the winner.  Returns (records [nhyp, RECORD], best); with on_device the
    winner stays a 0-d device tensor, so the step queues no host synchronisation of its own."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        allrec = torch.empty((world * per, RECORD), dtype=torch.float64, device=rec_local.device)
        dist.all_gather_into_tensor(allrec, rec_local.contiguous(), group=group)
    else:
        allrec = rec_local
    allrec = allrec[:nhyp]
    best = first_argmin_t(allrec[:, 0])
    return allrec, (best if on_device else int(best))


def icp_batch_sharded(local_fn, T0s, group=None, device="cpu"):
    """Shard the hypotheses T0s [H,4,4] over the ranks of `group`, run `local_fn(T0s_local)` ->
    dict(T [n,4,4], rmse, n_used, status) on each, all-gather, arg-min.  Every rank returns the full
    result dict (T, rmse, n_used, status, best)."""
    T0s = np.asarray(T0s, dtype=np.float64).reshape(-1, 4, 4)
    nhyp = T0s.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi, per = shard_range(nhyp, rank, world)
    if hi > lo:
        r = local_fn(T0s[lo:hi])
        rec = pack_records(r["T"], r["rmse"], r["n_used"], r["status"], per, device)
    else:
        rec = torch.full((per, RECORD), float("nan"), dtype=torch.float64, device=device)
    allrec, best = gather_and_pick(rec, nhyp, per, group)
    a = allrec.cpu().numpy()
    return dict(T=a[:, 1:17].reshape(nhyp, 4, 4).copy(), rmse=a[:, 0].copy(), n_used=a[:, 17].astype(np.int32),
                status=a[:, 18].astype(np.int32), best=best)
