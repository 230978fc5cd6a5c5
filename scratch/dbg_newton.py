import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from modelspec import *
import xara_b200 as xb
spec = brick_block(4, 4, 6, mat=J2_STEEL, lx=1.0, ly=1.0, lz=3.0, load=(1.2, 0.0, -0.5))
O = OracleBackend(spec, 1, 1); ptr, idx = O.csr(); neq=O.neq; ids=O.ids()
D = xb.DeviceModel.from_spec(spec, 1, 1).to_device(0)
solve = lambda A,B: spla.spsolve(sp.csr_matrix((A, idx, ptr), shape=(neq, neq)).tocsc(), B)
rel = lambda a,b: np.abs(a-b).max()/max(np.abs(b).max(),1e-300)
u = np.zeros((spec.nn,3)); lam=0
for step in range(4):
    lam += 1.0
    O.apply_load(lam); D.apply_load(lam)
    D.update()
    Bo = O.form_unbalance(); Bd = D.form_unbalance()
    print("step", step, "B0", rel(Bd,Bo))
    for it in range(10):
        Ao = O.form_tangent(); Ad = D.form_tangent()
        dU = solve(Ao, Bo)
        u[ids>=0] += dU[ids[ids>=0]]
        O.set_trial_disp(u); D.incr_trial_disp(dU); D.update()
        Bo = O.form_unbalance(); Bd = D.form_unbalance()
        print("  it", it, "|dU| %.2e"%np.linalg.norm(dU), "A", "%.1e"%rel(Ad,Ao), "B", "%.1e"%rel(Bd,Bo), "u", "%.1e"%rel(D.trial_disp(),u))
        if np.linalg.norm(dU) < 3e-8: break
    O.commit(); D.commit()
