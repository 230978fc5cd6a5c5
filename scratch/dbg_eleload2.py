import sys, numpy as np
sys.path[:0] = ['.', 'tests']
import xara_b200 as xb
from modelspec import *
def relerr(a, b): return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)
dim = 2
rng = np.random.default_rng(5)
spec = with_beam_gravity(frame2d(3, 3, 2), seed=3)
loaded = {t for t, *_ in spec.beam_loads}
O = OracleBackend(spec, 1, 0)
D = xb.DeviceModel.from_spec(spec, 1, 0).to_device(0)
ids = O.ids(); tags = O.fe_ids(6)[0]
sc = np.asarray((0.02, 0.02, 2e-4))
u = np.zeros((spec.nn, spec.ndf))
for s_ in range(3):
    if s_ != 2:
        u = u + rng.normal(0, 1.0, (spec.nn, spec.ndf)) * sc * 0.1; u[ids < 0] = 0
    lam = 0.25 * (s_ + 1)
    O.apply_load(lam); O.set_trial_disp(u)
    D.apply_load(lam); D.set_trial_disp(u); D.update()
    D.form_tangent(); D.form_unbalance(); O.form_tangent()
    if s_ == 2:
        for e in range(O.ne):
            print(e, int(tags[e]), "loaded" if int(tags[e]) in loaded else "column", "K", relerr(D.element_tangent(e, 6), O.ele_tangent(e, 6)), "R", relerr(D.element_resid(e, 6), O.ele_resid(e, 6)))
    O.commit(); D.commit()
