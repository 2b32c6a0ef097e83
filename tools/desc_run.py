"""GPU: the descriptor stage (getSpacialHistogramDescriptors.m) at a realistic size: keypoints on a 1 M-point model (~1500 points
per neighbourhood of R = 3.5), model with / without the uniform grid: python tools/desc_run.py [n_keypoints]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcreg_b200 as P
from pcreg_b200 import synth
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
P.init(0)
model = np.asarray(synth.make_model(1_000_000, 1003), dtype=np.float64)
g = synth.rng(5)
kp = model[g.integers(0, model.shape[0], nk)] + g.normal(0, 0.2, (nk, 3))
opts = dict(min_pts=500, max_pts=6000, R=3.5, thVar=(1.0, 1.0), k=0.85, ALIGN_POINTS=True)
res = {}
for name, m in (("grid", P.Model(model, grid=True, voxel_map=-1)), ("no grid", P.Model(model))):
    sub = kp if name == "grid" else kp[: max(1, nk // 10)]
    P.getSpacialHistogramDescriptors(m, sub[:64], opts)
    t0 = time.perf_counter()
    f, d = P.getSpacialHistogramDescriptors(m, sub, opts)
    dt = time.perf_counter() - t0
    print("%s: %d keypoints -> %d descriptors in %.1f ms (%.1f us per keypoint)" % (name, sub.shape[0], f.shape[0], dt * 1e3, dt * 1e6 / sub.shape[0]))
    res[name] = (f, d)
    m.destroy()
n = res["no grid"][0].shape[0]
print("same descriptors:", bool(np.array_equal(res["grid"][1][:n], res["no grid"][1]) and np.array_equal(res["grid"][0][:n], res["no grid"][0])))
