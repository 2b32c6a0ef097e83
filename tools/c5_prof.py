"""One C5-shaped ICP run for an ncu capture of the dense-model kernels: python tools/c5_prof.py [n_rot]"""
import sys
sys.path.insert(0, '/root/repo')
import pcreg_b200 as P
from pcreg_b200 import synth
P.init(0)
model = synth.make_model(16_000_000, 1005)
src, T_gt, c = synth.make_source(model[::16], 65536, 0.3, 1005)
m = P.Model(model, grid=True)
nrot = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T0 = synth.pose_grid(T_gt, c, nrot, (4, 4, 2), 10.0, 2.0, 7)
r = P.icp_batch(m, src, T0, mode=P.ICP_KNN, iters=20, nn=P.NN_GRID)
print('best rmse', r['rmse'][r['best']])
