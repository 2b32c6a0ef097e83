"""GPU: fixed per-call costs of the device-resident ICP call at small batch sizes (strong scaling): event time of the whole call vs
the fused kernel's own time, for H = 148 ... 4096 hypotheses of C3."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcreg_b200 as P
from pcreg_b200 import torch_ops
from bench import WORKLOADS, make_inputs
P.init(0)
w = WORKLOADS["c3"]
model, src, T0, w_src, T_gt = make_inputs(w, 0)
m = P.Model(model, grid=True)
dev = torch.device("cuda", 0)
opts = P.icp_opts(mode=P.ICP_KNN, iters=30, k_frac=0.85, R_w=3.5, nn=P.NN_GRID)
src_cm = torch_ops.src_to_abi_t(torch.from_numpy(src).to(dev))
for H in (148, 296, 512, 592, 1024, 2048, 4096):
    T0_abi = torch_ops.T_to_abi_t(torch.from_numpy(np.ascontiguousarray(T0[:: 4096 // H][:H])).to(dev))
    out = torch_ops.IcpDeviceBuffers(H, src.shape[0], 30, dev)
    for _ in range(3):
        torch_ops.icp_batch_device(m, src_cm, None, T0_abi, opts, out)
    torch.cuda.synchronize()
    ts, ws = [], []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        torch_ops.icp_batch_device(m, src_cm, None, T0_abi, opts, out)
        e1.record(); torch.cuda.synchronize()
        ws.append(1e3 * (time.perf_counter() - t0)); ts.append(e0.elapsed_time(e1))
    P.set_profiling(2)
    torch_ops.icp_batch_device(m, src_cm, None, T0_abi, opts, out); torch.cuda.synchronize()
    p = P.last_profile(); P.set_profiling(False)
    print("H %5d: call (events) %.3f ms, wall %.3f ms, fused kernel %.3f ms, fused %d, per-hyp-slot %.3f ms" % (
        H, np.median(ts), np.median(ws), p["nn_ms"], p["fused"], p["nn_ms"] / max(1, -(-H // 296))))
