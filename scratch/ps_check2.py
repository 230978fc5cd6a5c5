import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
rng=np.random.default_rng(4)
spec=quad_plane_stress_pressure(6,4,1,-2.0,mat=J2_STEEL,seed=35)
for numberer,soe in ((0,0),(1,1)):
    O=OracleBackend(spec,numberer,soe); R=RefBackend(spec,numberer,soe)
    print(np.abs(O.form_tangent()-R.form_tangent()).max()/np.abs(R.form_tangent()).max())
    ids=O.ids()
    for s in range(4):
        u=rng.normal(0,2e-3*(s+1),(spec.nn,2)); u[ids<0]=0
        O.set_trial_disp(u); R.set_trial_disp(u); O.apply_load(.3*s); R.apply_load(.3*s)
        A,Ar=O.form_tangent(),R.form_tangent(); B,Br=O.form_unbalance(),R.form_unbalance()
        print('  ',np.abs(A-Ar).max()/np.abs(Ar).max(), np.abs(B-Br).max()/np.abs(Br).max())
        if s%2==0: O.commit(); R.commit()
