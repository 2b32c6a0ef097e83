import sys
sys.path.insert(0, '/root/repo')
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
P.init(0)
w = WORKLOADS['c3']
model, src, T0, w_src, T_gt = make_inputs(w, 0)
m = P.Model(model, grid=True)
P.set_profiling(True)
r = P.icp_batch(m, src, T0, mode=P.ICP_KNN, iters=30, nn=P.NN_GRID)
p = P.last_profile()
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in p.items()})
print('answered from list %.3f; entries read per answered %.1f; gathered per answered %.1f; rowscan q %.3f; walked q %.3f' % (
    p['certified_queries'] / p['nn_queries'], p['list_entries_read'] / p['certified_queries'], p['list_points_gathered'] / p['certified_queries'],
    p['rowscan_queries'] / p['nn_queries'], p['walked_queries'] / p['nn_queries']))
