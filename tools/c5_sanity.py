"""C5-scale sanity (BASELINE.json configs[4]): 16M-point model, grid NN, 65k source points; compares grid vs
brute on the device (both exact) and times model creation.  Manual tool, not part of the test suite."""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import pcreg_b200 as P
from pcreg_b200 import synth
P.init(0)
t=time.time(); model=synth.make_model(16_000_000, 1005); print('synth', time.time()-t)
src,T_gt,c=synth.make_source(model[::16], 65536, 0.3, 1005)
t=time.time(); m=P.Model(model, grid=True); print('model_create s', time.time()-t, m.grid_info())
q=synth.apply_T(src, T_gt)
P.set_profiling(True)
t=time.time(); gi,gd=m.nn_search(q, P.NN_GRID); print('grid nn s', time.time()-t, P.last_profile()['nn_ms'])
t=time.time(); bi,bd=m.nn_search(q, P.NN_BRUTE); print('brute nn s', time.time()-t, P.last_profile()['nn_ms'], 'TF', 6*65536*16e6/P.last_profile()['nn_ms']/1e9)
print('grid==brute', np.array_equal(gi,bi), np.array_equal(gd,bd))
T0=synth.pose_grid(T_gt,c,4,(4,4,2),10.0,2.0,7)
t=time.time(); r=P.icp_batch(m,src,T0,mode=P.ICP_KNN,iters=20,nn=P.NN_GRID); print('icp 128 hyp x 65k x 20 it s', time.time()-t, P.last_profile()['nn_ms'], 'best rmse', r['rmse'][r['best']])
pr = P.last_profile()
nq = pr['nn_queries']
print('C5-shaped ICP: %.2f G queries/s on the NN kernels, list-answered %.3f, list %.1f ms rowscan %.1f ms walk %.1f ms update %.1f ms' % (
    nq / pr['nn_ms'] / 1e6, pr['certified_queries'] / nq, pr['list_ms'], pr['rowscan_ms'], pr['walk_ms'], pr['update_ms']))
T0 = synth.pose_grid(T_gt, c, 16, (8, 4, 4), 10.0, 2.0, 7)          # 2048 hypotheses = one GPU's share of C5 (16k over 8 GPUs)
t = time.time(); r = P.icp_batch(m, src, T0, mode=P.ICP_KNN, iters=20, nn=P.NN_GRID); dt = time.time() - t
pr = P.last_profile(); nq = pr['nn_queries']
print('C5 per-GPU share: 2048 hyp x 65536 x 21 passes in %.2f s wall = %.2f G queries/s (NN kernels %.0f ms, list-answered %.3f)' % (
    dt, nq / dt / 1e9, pr['nn_ms'], pr['certified_queries'] / nq))
