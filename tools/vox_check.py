"""GPU: build the Voronoi voxel map of a C3-sized model, print its facts, time raw NN passes of both grid paths."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcreg_b200 as P
from pcreg_b200 import synth

nm = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
P.init(0)
model = synth.make_model(nm, 1003)
t0 = time.time()
m = P.Model(model, grid=True)
print("model create %.2f s" % (time.time() - t0), m.grid_info(), m.voxel_info(), flush=True)
src, T_gt, c = synth.make_source(model, 5000, 0.3, 1003)
T0 = synth.pose_grid(T_gt, c, 16, (8, 8, 4), 20.0, 2.0, 1003)[:: max(1, 4096 // 64)][:64]
q = np.concatenate([synth.apply_T(src, T) for T in T0])
for kind, name in ((P.NN_GRID, "grid"),):
    P.set_profiling(True)
    i1, d1 = m.nn_search(q, kind)
    pr = P.last_profile()
    P.set_profiling(False)
    print(name, "first-pass-like queries: %d in %.3f ms = %.1f Mq/s" % (q.shape[0], pr["nn_ms"], q.shape[0] / pr["nn_ms"] / 1e3),
          {k: pr[k] for k in ("certified_queries", "walked_queries", "list_entries_read", "list_points_gathered", "voxel_map")}, flush=True)
m2 = P.Model(model, grid=True, voxel_map=-1)
P.set_profiling(True)
i2, d2 = m2.nn_search(q, P.NN_GRID)
pr = P.last_profile()
P.set_profiling(False)
print("old grid path: %.3f ms" % pr["nn_ms"], "identical:", bool(np.array_equal(i1, i2) and np.array_equal(d1, d2)), flush=True)
