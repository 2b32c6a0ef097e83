"""GPU: the C4 polish (PLAIN + thDist2, 20k source vs 2M model) on a reduced batch: kernel ms and phase shares."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcreg_b200 as P
from pcreg_b200 import synth
from bench import C4 as c
H = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P.init(0)
model = synth.make_model(c["nm"], c["seed"])
src, T_gt, ctr = synth.make_source(model, c["ns"], 0.3, c["seed"])
g = synth.rng(5)
T0 = np.stack([synth.perturb_pose(T_gt, ctr, synth.rot_axis_angle(g.standard_normal(3), np.deg2rad(g.uniform(0, 6))), g.normal(0, 1.0, 3)) for _ in range(H)])
m = P.Model(model, grid=True, max_voxels=int(os.environ.get("TOOL_MAX_VOXELS", "0")), voxel_scale=float(os.environ.get("TOOL_VOX_SCALE", "0")))
print(m.voxel_info())
for prof in (2, 1):
    P.set_profiling(prof)
    for _ in range(2):
        r = P.icp_batch(m, src, T0, mode=P.ICP_PLAIN, iters=c["iters"], thDist2=c["thDist2"], nn=P.NN_GRID)
    p = P.last_profile()
    print("entries/query", p.get("counters", {}) if isinstance(p.get("counters"), dict) else "", end=" ")
    print("profiling", prof, "H", H, "kernel ms %.1f" % p["nn_ms"], "fused", p["fused"], p["fused_phase_share"], "rmse %.17g" % r["rmse"][r["best"]])
P.set_profiling(False)
