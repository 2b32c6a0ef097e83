import sys, numpy as np
sys.path.insert(0,'/root/repo')
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
P.init(0)
w=WORKLOADS['c3']
model,src,T0,w_src,T_gt=make_inputs(w,0)
m=P.Model(model,grid=True)
P.set_profiling(True)
for it in (0,1):
    r=P.icp_batch(m,src,T0[:1024],mode=P.ICP_KNN,iters=it,nn=P.NN_GRID)
    pr=P.last_profile()
    nq=pr['nn_queries']
    print(it, 'queries', nq, 'pts/q', pr['grid_points_visited']/nq, 'leaf/q', pr['grid_cells_visited']/nq, 'nodes/q', pr['grid_nodes_popped']/nq, 'nn_ms', pr['nn_ms'])
