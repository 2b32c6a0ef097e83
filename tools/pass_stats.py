"""Per-pass grid-NN statistics on C3 (cumulative counters differenced over the iteration count)."""
import sys
sys.path.insert(0, '/root/repo')
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
P.init(0)
w = WORKLOADS['c3']
model, src, T0, w_src, T_gt = make_inputs(w, 0)
m = P.Model(model, grid=True)
P.set_profiling(True)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prev = None
for it in (0, 1, 2, 3, 4, 8, 16, 30):
    for rep in range(2):
        r = P.icp_batch(m, src, T0[:H], mode=P.ICP_KNN, iters=it, nn=P.NN_GRID)
    pr = P.last_profile()
    cur = (it + 1, pr['grid_points_visited'], pr['grid_cells_visited'], pr['grid_nodes_popped'], pr['nn_ms'], pr['certified_queries'], pr['walked_queries'], pr['rowscan_queries'])
    if prev is None:
        d = cur
    else:
        d = tuple(c - p for c, p in zip(cur, prev))
    nq = d[0] * H * src.shape[0]
    print('passes %d..%d: pts/q %.1f rows/q %.2f nodes/q %.2f  ms/pass %.3f  Mq/s %.0f  from-list %.3f walked %.4f rowscan %.4f' % (
        (prev[0] if prev else 0), it, d[1] / nq, d[2] / nq, d[3] / nq, d[4] / d[0], nq / d[4] / 1e3, d[5] / nq, d[6] / nq, d[7] / nq))
    prev = cur
