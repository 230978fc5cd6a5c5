import sys, faulthandler; faulthandler.enable()
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
mk = lambda: brick_block(4, 4, 6, mat=J2_STEEL, lx=1.0, ly=1.0, lz=3.0, load=(1.2, 0.0, -0.5))
C = RefBackend(mk(), 1, 0, dlambda=1/8, test=0, tol=1e-8, max_iter=3)
rc, iters, norms = C.analyze_static(8)
print('rc', rc, iters.tolist())
u=C.get_trial_disp(); print(np.abs(u).max())
