import sys
sys.path.insert(0, '/root/repo')
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
P.init(0)
w = WORKLOADS['c2g']
model, src, T0, w_src, T_gt = make_inputs(w, 0)
m = P.Model(model, grid=True)
for rep in range(2):
    r = P.icp_batch(m, src, T0, mode=P.ICP_WEIGHTED, iters=100, nn=P.NN_GRID, w_src=w_src)
print(r['rmse'])
