"""C5-shaped work counters (16M-point model, 65 536 source points): points / rows / nodes per query and the kernel split.
python tools/c5_stats.py [n_rot] -> n_rot x 32 hypotheses."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import pcreg_b200 as P
from pcreg_b200 import synth
P.init(0)
model = synth.make_model(16_000_000, 1005)
src, T_gt, c = synth.make_source(model[::16], 65536, 0.3, 1005)
m = P.Model(model, grid=True)
print(m.grid_info())
P.set_profiling(True)
nrot = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T0 = synth.pose_grid(T_gt, c, nrot, (4, 4, 2), 10.0, 2.0, 7)
prev = None
for it in (0, 1, 2, 4, 8, 20):
    for rep in range(2):
        r = P.icp_batch(m, src, T0, mode=P.ICP_KNN, iters=it, nn=P.NN_GRID)
    pr = P.last_profile()
    cur = np.array([it + 1, pr['grid_points_visited'], pr['grid_cells_visited'], pr['grid_nodes_popped'], pr['nn_ms'], pr['certified_queries'],
                    pr['walked_queries'], pr['rowscan_queries'], pr['list_ms'], pr['rowscan_ms'], pr['walk_ms'], pr['update_ms']], dtype=float)
    d = cur if prev is None else cur - prev
    nq = d[0] * T0.shape[0] * src.shape[0]
    print('passes %d..%d: pts/q %.1f rows/q %.2f nodes/q %.2f ms/pass %.2f Mq/s %.0f from-list %.3f walked %.4f rowscan %.4f | list %.1f rowscan %.1f walk %.1f update %.1f ms' % (
        0 if prev is None else prev[0], it, d[1] / nq, d[2] / nq, d[3] / nq, d[4] / d[0], nq / d[4] / 1e3, d[5] / nq, d[6] / nq, d[7] / nq, d[8], d[9], d[10], d[11]))
    prev = cur
print({k: v for k, v in pr.items() if 'list' in k or 'gather' in k or 'entries' in k})
