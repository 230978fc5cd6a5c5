import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
rng=np.random.default_rng(2)
cases=[(brick_block(3,3,3,distort=0.1),4e-3),(quad_plane(5,4,mat=J2_STEEL,lx=5.,ly=4.,distort=.2),3e-3),(quad_plane_stress_pressure(5,4,1,1.5,mat=J2_STEEL),3e-3),
       (quad_plane_stress_pressure(5,4,1,1.5),2e-2),(frame2d(2,2,2),np.array((0.006,0.003,6e-5))),(frame3d(1,1,2),np.array((0.015,0.015,0.003,1e-4,1e-4,1e-4)))]
for spec,sc in cases:
    O=OracleBackend(spec,1,1); R=RefBackend(spec,1,1); ids=O.ids()
    A0=O.form_tangent().copy()
    for s in range(3):
        u=rng.normal(0,1,(spec.nn,spec.ndf))*sc*(s+1); u[ids<0]=0
        for m in (O,R): m.set_trial_disp(u); m.apply_load(.4*(s+1)); m.commit()
    for m in (O,R): m.revert_to_start()
    A,Ar=O.form_tangent(),R.form_tangent(); B,Br=O.form_unbalance(),R.form_unbalance()
    print('after reset', np.abs(A-Ar).max()/np.abs(Ar).max(), np.abs(B-Br).max()/max(np.abs(Br).max(),1e-300), 'back to initial:', np.abs(A-A0).max()/np.abs(A0).max())
    u=rng.normal(0,1,(spec.nn,spec.ndf))*sc; u[ids<0]=0
    for m in (O,R): m.set_trial_disp(u); m.apply_load(.3)
    A,Ar=O.form_tangent(),R.form_tangent(); B,Br=O.form_unbalance(),R.form_unbalance()
    print('   next step', np.abs(A-Ar).max()/np.abs(Ar).max(), np.abs(B-Br).max()/np.abs(Br).max())
