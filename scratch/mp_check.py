import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
from golden_cases import CASES
import xara_b200 as xb
for name,(mk,numberer,soe,scale) in CASES.items():
    if 'equaldof' not in name: continue
    spec=mk()
    O=OracleBackend(spec,numberer,soe); R=RefBackend(spec,numberer,soe)
    print(name, O.neq,R.neq,O.nnz,R.nnz, np.array_equal(O.ids(),R.ids()), all(np.array_equal(a,b) for a,b in zip(O.csr(),R.csr())))
    D=xb.DeviceModel.from_spec(spec,numberer,soe)
    print('  host', D.neq, D.nnz, np.array_equal(D.ids(),R.ids()), all(np.array_equal(a,b) for a,b in zip(D.pattern(),R.csr())))
    rng=np.random.default_rng(1)
    for s in range(3):
        u=rng.normal(0,1,(spec.nn,spec.ndf))*np.asarray(scale)*(s+1); u[O.ids()<0]=0; tie(spec,u)
        O.set_trial_disp(u); R.set_trial_disp(u); O.apply_load(0.3*(s+1)); R.apply_load(0.3*(s+1))
        A,Ar=O.form_tangent(),R.form_tangent(); B,Br=O.form_unbalance(),R.form_unbalance()
        print('  ', np.abs(A-Ar).max()/np.abs(Ar).max(), np.abs(B-Br).max()/np.abs(Br).max())
        O.commit(); R.commit()
