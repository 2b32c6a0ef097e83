"""C4-shaped sanity (BASELINE.json configs[3], scaled): RANSAC-seeded hypotheses polished by a 20-iteration ICP,
20k source vs 2M model, grid NN, PLAIN mode with thDist2 = 4.  Checks grid == brute on a subset and prints the rate."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import pcreg_b200 as P
from pcreg_b200 import synth
P.init(0)
model = synth.make_model(2_000_000, 1004)
src, T_gt, c = synth.make_source(model, 20000, 0.3, 1004)
m = P.Model(model, grid=True)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
# hypotheses as a RANSAC stage would seed them: the true pose perturbed by a few degrees / mm
g = synth.rng(5)
T0 = np.stack([synth.perturb_pose(T_gt, c, synth.rot_axis_angle(g.standard_normal(3), np.deg2rad(g.uniform(0, 6))), g.normal(0, 1.0, 3)) for _ in range(H)])
for rep in range(2):
    t = time.time()
    r = P.icp_batch(m, src, T0, mode=P.ICP_PLAIN, iters=20, thDist2=4.0, nn=P.NN_GRID, return_idx=True)
    dt = time.time() - t
print('%d hyp x 20k x 21 passes: %.3f s = %.2f G queries/s, best rmse %.4f' % (H, dt, H * 20000 * 21 / dt / 1e9, r['rmse'][r['best']]))
b = P.icp_batch(m, src, T0[:4], mode=P.ICP_PLAIN, iters=20, thDist2=4.0, nn=P.NN_BRUTE, return_idx=True)
print('grid == brute on 4 hypotheses:', np.array_equal(r['idx'][:4], b['idx']), np.array_equal(r['T'][:4], b['T']), np.array_equal(r['rmse'][:4], b['rmse']))
