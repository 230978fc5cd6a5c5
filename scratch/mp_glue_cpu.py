import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
def mk():
    sp = soil_column_equaldof(12, mat=J2_STEEL, distort=0.1)
    sp.loads = np.array([[1 + 2 * 12, 22.0 * 8, -3.0 * 8], [1 + 2 * 6, 10.0 * 8, 0.0]]); return sp
for numberer,soe in ((1,0),(0,1)):
  for tol in (1e-6,1e-7,1e-8,1e-9):
    C = RefBackend(mk(), numberer, soe, dlambda=1/8, test=0, tol=tol, max_iter=25)
    rc, iters, norms = C.analyze_static(8)
    margin = min(min(norms[s, iters[s] - 2] / tol if iters[s] > 1 else 1e9, tol / max(norms[s, iters[s] - 1], 1e-300)) for s in range(8))
    print(numberer,soe,tol,rc,iters.tolist(),margin)
