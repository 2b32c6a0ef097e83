"""GPU: C5-shaped work (16M-point model, 65 536 source points) on the band-limited voxel map vs the grid kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P.init(0)
w = dict(WORKLOADS["c5"]); w["hyp"] = 2048
model, src, T0, w_src, T_gt = make_inputs(w, 0)
T0 = T0[:: max(1, T0.shape[0] // H)][:H]
t0 = time.time()
m = P.Model(model, grid=True)
print("model create %.1f s" % (time.time() - t0), m.grid_info(), m.voxel_info(), flush=True)
res = {}
for name, mm in (("voxel map", m), ("grid kernels", None)):
    if mm is None:
        mm = P.Model(model, grid=True, voxel_map=-1)
    for prof in (2, 1):
        P.set_profiling(prof)
        t0 = time.time()
        r = P.icp_batch(mm, src, T0, mode=P.ICP_KNN, iters=20, nn=P.NN_GRID, return_idx=(prof == 1))
        dt = time.time() - t0
        p = P.last_profile()
        print(name, "profiling", prof, "wall %.2f s nn %.1f ms (list %.1f rows %.1f walk %.1f) update %.1f ms | answered %.3f walked %.3f entries/q %.1f" % (
            dt, p["nn_ms"], p["list_ms"], p["rowscan_ms"], p["walk_ms"], p["update_ms"], p["certified_queries"] / p["nn_queries"],
            p["walked_queries"] / p["nn_queries"], p["list_entries_read"] / max(1.0, p["certified_queries"])), flush=True)
    P.set_profiling(False)
    res[name] = r
a, b = res["voxel map"], res["grid kernels"]
print("identical:", all(np.array_equal(a[k], b[k]) for k in ("T", "rmse", "idx", "n_used")))
