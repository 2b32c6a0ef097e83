"""Per-step wall / event times of the C3 device path (host hiccup diagnosis): python tools/step_times.py [nsteps]"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import pcreg_b200 as P
from pcreg_b200 import torch_ops
from bench import WORKLOADS, make_inputs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
P.init(0)
w = WORKLOADS['c3']
model, src, T0, w_src, T_gt = make_inputs(w, 0)
m = P.Model(model, grid=True)
dev = torch.device('cuda', 0)
opts = P.icp_opts(mode=P.ICP_KNN, iters=30, k_frac=0.85, R_w=3.5, nn=P.NN_GRID)
src_cm = torch_ops.src_to_abi_t(torch.from_numpy(src).to(dev))
T0_abi = torch_ops.T_to_abi_t(torch.from_numpy(T0).to(dev))
out = torch_ops.IcpDeviceBuffers(T0.shape[0], src.shape[0], 30, dev)
ts = []
for i in range(n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    torch_ops.icp_batch_device(m, src_cm, None, T0_abi, opts, out)
    e1.record(); torch.cuda.synchronize()
    ts.append((1e3 * (time.perf_counter() - t0), e0.elapsed_time(e1)))
print(' '.join('%.0f/%.0f' % t for t in ts))
print('median wall %.1f  min %.1f' % (np.median([t[0] for t in ts]), min(t[0] for t in ts)))
