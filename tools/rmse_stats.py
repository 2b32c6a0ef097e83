import sys, numpy as np
sys.path.insert(0,'/root/repo')
import pcreg_b200 as P
from bench import WORKLOADS, make_inputs
P.init(0)
w=WORKLOADS['c3']
model,src,T0,w_src,T_gt=make_inputs(w,0)
m=P.Model(model,grid=True)
for it in (0,1,2,5,10,30):
    r=P.icp_batch(m,src,T0,mode=P.ICP_KNN,iters=it,nn=P.NN_GRID)
    q=np.percentile(r['rmse'],[0,10,25,50,75,90,100])
    print(it, np.round(q,3))
