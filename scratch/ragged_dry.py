import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
from test_host_setup import ragged_spec
for seed in range(6):
    rng = np.random.default_rng(500 + seed); spec = ragged_spec(seed)
    O = OracleBackend(spec, 1, 1); ids = O.ids()
    for s in range(3):
        u = rng.normal(0, 2e-3 * (s + 1), (spec.nn, 3)); u[ids < 0] = 0; tie(spec, u)
        O.set_trial_disp(u); O.apply_load(0.3 * s); A=O.form_tangent(); B=O.form_unbalance(); O.commit()
    print(seed, spec.ne, O.neq, len(spec.equal_dofs), np.abs(A).max(), np.abs(B).max())
