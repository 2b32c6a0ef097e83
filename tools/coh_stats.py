import sys, numpy as np
sys.path.insert(0,'/root/repo')
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
P.init(0)
w=WORKLOADS['c3']
model,src,T0,w_src,T_gt=make_inputs(w,0)
m=P.Model(model,grid=True)
P.set_profiling(True)
for it in (2,5,10,30):
    r=P.icp_batch(m,src,T0[:1024],mode=P.ICP_KNN,iters=it,nn=P.NN_GRID)
    pr=P.last_profile()
    print(it, 'queries', pr['nn_queries'], 'certified', pr['certified_queries'], 'frac', pr['certified_queries']/pr['nn_queries'], 'pts/q', pr['grid_points_visited']/pr['nn_queries'], 'rows/q', pr['grid_cells_visited']/pr['nn_queries'], 'nn_ms', pr['nn_ms'])
