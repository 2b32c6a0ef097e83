"""Single-GPU timing of the C3 shares that each rank of an N-GPU run would get (data-dependent work balance)."""
import sys, time
sys.path.insert(0, '/root/repo')
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
P.init(0)
w = WORKLOADS['c3']
m = None
for rank, world in [(0, 1), (0, 2), (1, 2), (0, 8), (3, 8), (7, 8)]:
    model, src, T0, w_src, T_gt = make_inputs(w, rank, world)
    if m is None:
        m = P.Model(model, grid=True)
    for rep in range(3):
        t0 = time.perf_counter()
        r = P.icp_batch(m, src, T0, mode=P.ICP_KNN, iters=30, nn=P.NN_GRID)
        dt = time.perf_counter() - t0
    print('rank %d of %d: %.1f ms  best rmse %.5f' % (rank, world, dt * 1e3, r['rmse'][r['best']]))
