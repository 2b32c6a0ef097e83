"""GPU: per-level build times of the voxel map (PCREG_DEBUG_HOST=1): python tools/vox_build_times.py c3|c5"""
import sys, os, time
os.environ["PCREG_DEBUG_HOST"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcreg_b200 as P
from pcreg_b200 import synth
nm, seed = dict(c3=(1_000_000, 1003), c4=(2_000_000, 1004), c5=(16_000_000, 1005))[sys.argv[1]]
P.init(0)
model = synth.make_model(nm, seed)
t0 = time.time()
m = P.Model(model, grid=True)
print("create %.2f s" % (time.time() - t0), m.voxel_info())
