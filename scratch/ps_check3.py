import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
rng=np.random.default_rng(9)
spec=quad_plane_stress_pressure(7,5,1,1.5,mat=J2_STEEL,seed=37)
O=OracleBackend(spec,1,0); R=RefBackend(spec,1,0); ids=O.ids()
def cmp(tag):
    A,Ar=O.form_tangent(),R.form_tangent(); B,Br=O.form_unbalance(),R.form_unbalance()
    print(tag, np.abs(A-Ar).max()/np.abs(Ar).max(), np.abs(B-Br).max()/np.abs(Br).max())
for s in range(6):
    u=rng.normal(0,2e-3*(s+1),(spec.nn,2)); u[ids<0]=0
    O.set_trial_disp(u); R.set_trial_disp(u); O.apply_load(.2*s); R.apply_load(.2*s); cmp(f'step{s}')
    if s in (1,4): O.commit(); R.commit()
    if s==3: O.revert(); R.revert(); cmp('after revert')
