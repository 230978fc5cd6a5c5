import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
rng=np.random.default_rng(0)
for spec in (brick_block(3,3,3,distort=0.1), frame2d(2,2,2)):
    O=OracleBackend(spec,0,1); R=RefBackend(spec,0,1); ids=O.ids()
    sc = 4e-3 if spec.ndf==3 and spec.ndm==3 else np.array((0.006,0.003,6e-5))
    u1=rng.normal(0,1,(spec.nn,spec.ndf))*sc; u1[ids<0]=0
    for m in (O,R): m.set_trial_disp(u1); m.apply_load(0.5); m.commit()
    u2=u1+rng.normal(0,1,(spec.nn,spec.ndf))*sc; u2[ids<0]=0
    for m in (O,R): m.set_trial_disp(u2); m.apply_load(0.9); m.revert()
    B,Br=O.form_unbalance(),R.form_unbalance(); A,Ar=O.form_tangent(),R.form_tangent()
    print(np.abs(B-Br).max()/np.abs(Br).max(), np.abs(A-Ar).max()/np.abs(Ar).max())
