import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from modelspec import *
import xara_b200 as xb
from test_gpu_parity import _partitioned_pass, relerr
rng = np.random.default_rng(5)
mk = lambda: brick_block(6, 5, 7, mat=J2_STEEL, distort=0.2, seed=9, body=(0.0, 0.01, -0.02))
spec = mk()
G = xb.DeviceModel.from_spec(spec, 0, 0).to_device(0)
O = OracleBackend(spec, 0, 0)
ids = G.ids()
u = rng.normal(0, 2e-3, (spec.nn, 3)); u[ids < 0] = 0
G.set_trial_disp(u); G.update(); G.apply_load(0.0)
Ag, Bg = G.form_tangent(), G.form_unbalance()
O.set_trial_disp(u); O.apply_load(0.0)
print("single vs oracle", relerr(Ag, O.form_tangent()), relerr(Bg, O.form_unbalance()))
ranks = [xb.DeviceModel.from_spec(mk(), 0, 0, 2, r).to_device(0) for r in range(2)]
G.set_trial_disp(u); G.update()
Ag2, Bg2 = G.form_tangent(), G.form_unbalance()
print("after creating ranks", relerr(Ag2, Ag), relerr(Bg2, Bg))
res = _partitioned_pass(ranks, u, 0.0)
gptr,_ = G.pattern()
for m,(A,B) in zip(ranks,res):
    rows = m.row_eqns(); ptr,_ = m.pattern()
    Aref = np.concatenate([Ag[gptr[q]:gptr[q + 1]] for q in rows])
    print("rank", m.rank, "B eq", np.array_equal(B, Bg[rows]), relerr(B, Bg[rows]), "A eq", np.array_equal(A, Aref), relerr(A, Aref), m.peers())
