import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
which=sys.argv[1]
mat = J2_STEEL if sys.argv[2]=='j2' else ELASTIC
spec={'soil':lambda: soil_column_equaldof(5, mat=mat),'brick':lambda: brick_periodic_equaldof(2,2,2,mat=mat),'frame':lambda: frame2d_diaphragm_equaldof(2,2,1)}[which]()
for numberer in (0,1):
  for soe in (0,1):
    R=RefBackend(spec,numberer,soe)
    u=np.zeros((spec.nn,spec.ndf)); R.set_trial_disp(u); print(which, numberer, soe, 'ok', R.neq)
