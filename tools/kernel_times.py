"""One C3-shaped ICP call (H hypotheses, IT iterations) for an ncu launch list: python tools/kernel_times.py [H] [IT]"""
import sys
sys.path.insert(0, '/root/repo')
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 30
P.init(0)
w = WORKLOADS['c3']
model, src, T0, w_src, T_gt = make_inputs(w, 0)
m = P.Model(model, grid=True)
r = P.icp_batch(m, src, T0[:H], mode=P.ICP_KNN, iters=IT, nn=P.NN_GRID)
print('best rmse', r['rmse'][r['best']])
