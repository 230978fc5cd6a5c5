import sys, numpy as np
sys.path[:0] = ['.', 'tests']
import xara_b200 as xb
from modelspec import *
def relerr(a, b): return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)
import itertools
for dim, l2 in itertools.product((2, 3), (0.75, 0.8, 0.6, 1.5)):
    rng = np.random.default_rng(5)
    spec = with_beam_gravity(frame2d(3, 3, 2) if dim == 2 else frame3d(2, 1, 2), seed=3)
    O = OracleBackend(spec, 1, 0)
    D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
    ids = O.ids()
    sc = np.asarray((0.02, 0.02, 2e-4) if dim == 2 else (0.015, 0.015, 0.003, 1e-4, 1e-4, 1e-4))
    u = np.zeros((spec.nn, spec.ndf))
    for s_ in range(4):
        if s_ != 2:
            u = u + rng.normal(0, 1.0, (spec.nn, spec.ndf)) * sc * 0.1; u[ids < 0] = 0
        lam = 0.25 * (s_ + 1) if s_ != 2 else l2
        O.apply_load(lam); rc = O.set_trial_disp(u)
        D.apply_load(lam); D.set_trial_disp(u); D.update()
        print(dim, l2, s_, "orc rc", rc, "A", relerr(D.form_tangent(), O.form_tangent()), "B", relerr(D.form_unbalance(), O.form_unbalance()), flush=True)
        O.commit(); D.commit()
