import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
rng=np.random.default_rng(3)
kind,p=J2_STEEL
n=200
strains=np.cumsum(rng.normal(0,6e-4,(n,3)),axis=0)
commit=(rng.random(n)<0.6).astype(np.int32)
so,to=oracle_nd_path(kind,p,ND_PLANE_STRESS,strains,commit)
sr,tr=ref_nd_path(kind,p,ND_PLANE_STRESS,strains,commit)
print('stress err',np.abs(so-sr).max()/np.abs(sr).max(),'tangent err',np.abs(to-tr).max()/np.abs(tr).max(), 'plastic?', len(np.unique(np.round(tr[:,0,0],1))))
