"""GPU: the getLocalPoints count pass at the descriptor stage's full size (10^5 keypoints x 16 M-point model, R = 3.5) for an
ncu capture of k_local_grid: python tools/local_points_run.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcreg_b200 as P
from pcreg_b200 import synth, _lib as L
P.init(0)
model = np.asarray(synth.make_model(16_000_000, 1005), dtype=np.float64)
g = synth.rng(77)
kp = np.asfortranarray(model[g.integers(0, model.shape[0], 100_000)] + g.normal(0, 0.3, (100_000, 3)))
m = P.Model(model, grid=True, voxel_map=-1)
counts = np.empty(kp.shape[0], dtype=np.int64); status = np.empty(kp.shape[0], dtype=np.int32)
lib = L.lib()
for rep in range(3):
    t0 = time.perf_counter()
    L.check(lib.pcreg_local_points_count(m.handle, kp.ctypes.data_as(L.c_f64p), kp.shape[0], kp.shape[0], 3.5, 30, 6000,
                                         counts.ctypes.data_as(L.c_i64p), status.ctypes.data_as(L.c_i32p)), "count")
    print("count pass %.2f ms, %.3g points inside the balls" % ((time.perf_counter() - t0) * 1e3, float(counts.sum())))
