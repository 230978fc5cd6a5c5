import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
mat=J2_STEEL
for numberer in (0,1):
  for soe in (0,1):
    rng = np.random.default_rng(7)
    specs = [brick_block(3, 2, 2, mat=mat, distort=0.2, seed=11), quad_plane(5, 4, mat=mat, distort=0.2, seed=12)]
    specs += [soil_column_equaldof(5, mat=mat), brick_periodic_equaldof(2, 2, 2, mat=mat)]
    specs += [frame2d(2,2,2), frame3d(1,1,2), frame2d_diaphragm_equaldof(2,2,1)]
    for k,spec in enumerate(specs):
        beam = spec.groups[0].kind in (2, 3)
        O, R = OracleBackend(spec, numberer, soe), RefBackend(spec, numberer, soe)
        for s in range(3):
            sc = (0.02, 0.02, 2e-4) if spec.groups[0].kind == 2 else ((0.015, 0.015, 0.003, 1e-4, 1e-4, 1e-4) if beam else 2e-3)
            u = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * np.asarray(sc) * (s + 1); u[O.ids() < 0] = 0
            tie(spec, u)
            print(numberer, soe, k, s, flush=True)
            O.set_trial_disp(u); R.set_trial_disp(u)
            O.apply_load(0.3 * s); R.apply_load(0.3 * s)
            A,Ar=O.form_tangent(), R.form_tangent()
            print('   ', np.abs(A-Ar).max()/np.abs(Ar).max(), flush=True)
            O.commit(); R.commit()
