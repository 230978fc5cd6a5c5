# dry run of the oracle half of test_device_vs_oracle_load_history_equaldof (no device here)
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from modelspec import *
import xara_b200 as xb
for shape in ["soilcolumn", "brick", "brick_elastic", "frame2d"]:
  for numberer,soe in [(0,0),(1,1),(1,0)]:
    rng = np.random.default_rng(43)
    if shape == "soilcolumn":
        spec, sc = soil_column_equaldof(40, distort=0.2, seed=31), 2e-3
    elif shape == "brick":
        spec, sc = brick_periodic_equaldof(5, 4, 3, seed=32), 1.5e-3
    elif shape == "brick_elastic":
        spec, sc = brick_periodic_equaldof(3, 5, 4, mat=ELASTIC, dofs=(0, 1, 2), seed=33), 1.5e-3
    else:
        spec, sc = frame2d_diaphragm_equaldof(3, 4, 2), np.array((0.006, 0.003, 6e-5))
    mass = rng.uniform(0.01, 0.1, (spec.nn, spec.ndf))
    O = OracleBackend(spec, numberer, soe); O.set_mass(spec.node_tags, mass)
    D = xb.DeviceModel.from_spec(spec, setup=False); D.set_mass(spec.node_tags, mass); D.setup(numberer, soe)
    ids = O.ids()
    assert np.array_equal(D.ids(), ids)
    print(shape, numberer, soe, (np.bincount(ids[ids >= 0]) > 1).sum(), D.neq)
    for s in range(5):
        u = rng.normal(0, 1.0, (spec.nn, spec.ndf)) * sc * (s + 1); u[ids < 0] = 0
        tie(spec, u)
        O.set_trial_disp(u); O.apply_load(0.2*s)
        if s == 3:
            v, a = rng.normal(0, 1.0, (2, spec.nn, spec.ndf)); v[ids < 0] = 0; a[ids < 0] = 0
            tie(spec, v); tie(spec, a)
            O.set_rayleigh(0.3, 0.0, 0.0, 0.0); O.set_transient(1.0, 50.0, 1.0e4); O.set_vel_accel(v, a)
        A=O.form_tangent(); B=O.form_unbalance()
        assert np.isfinite(A).all() and np.isfinite(B).all()
        if s%2==0: O.commit()
