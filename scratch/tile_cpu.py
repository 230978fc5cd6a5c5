import sys, os, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
import xara_b200 as xb
spec = brick_block(48, 40, 36, distort=0.2, seed=3)
# shuffle element input order and tags to make it non-trivial
rng=np.random.default_rng(0); g=spec.groups[0]; p=rng.permutation(len(g.tags))
g.tags=g.tags[p]; g.conn=g.conn[p]; g.mat=g.mat[p]; g.par=g.par[p]
os.environ['XB_TILE']='0'; t=time.time(); D0=xb.DeviceModel.from_spec(spec,1,0); print('plain',time.time()-t)
os.environ['XB_TILE']='9472'; t=time.time(); D1=xb.DeviceModel.from_spec(spec,1,0); print('tiled',time.time()-t)
print(np.array_equal(D0.ids(),D1.ids()), all(np.array_equal(a,b) for a,b in zip(D0.pattern(),D1.pattern())), np.array_equal(D0.element_tags(),D1.element_tags()))
s0=D0.scatter_map(0,2000,24); s1=D1.scatter_map(0,2000,24); print(np.array_equal(s0,s1))
s0=D0.scatter_map(60000,61000,24); s1=D1.scatter_map(60000,61000,24); print(np.array_equal(s0,s1))
