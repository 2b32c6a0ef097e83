"""GPU: phase shares of the fused ICP kernel on C3 (clock64 of thread 0 per block)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
P.init(0)
w = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"])
if len(sys.argv) > 2: w["hyp"] = int(sys.argv[2])
model, src, T0, w_src, T_gt = make_inputs(w, 0)
m = P.Model(model, grid=True)
mode = dict(plain=P.ICP_PLAIN, knn=P.ICP_KNN, weighted=P.ICP_WEIGHTED)[w["mode"]]
for prof in (2, 1):
    P.set_profiling(prof)
    for _ in range(2):
        r = P.icp_batch(m, src, T0, mode=mode, iters=w["iters"], nn=P.NN_GRID, w_src=w_src)
    p = P.last_profile()
    print("profiling", prof, "kernel ms %.2f" % p["nn_ms"], "fused", p["fused"], p["fused_phase_share"],
          "entries/query %.2f gathers/query %.4f walked %d" % (p["list_entries_read"] / p["nn_queries"], p["list_points_gathered"] / p["nn_queries"], p["walked_queries"]))
P.set_profiling(False)
