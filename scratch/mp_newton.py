import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from modelspec import *
sys.argv=[sys.argv[0]]
import importlib.util
spec_ = importlib.util.spec_from_file_location("tg", "/root/repo/tests/test_gpu_parity.py")
tg = importlib.util.module_from_spec(spec_); spec_.loader.exec_module(tg)
def run(spec, nsteps):
    O = OracleBackend(spec, 1, 1); ptr, idx = O.csr(); neq = O.neq
    solve = lambda A,B: spla.spsolve(sp.csr_matrix((A, idx, ptr), shape=(neq, neq)).tocsc(), B)
    for tol in (1e-6, 3e-7, 1e-7, 3e-8, 1e-8, 3e-9, 1e-9):
        O = OracleBackend(spec, 1, 1); O._u = np.zeros((spec.nn, spec.ndf))
        h = tg._newton(O, solve, nsteps, 1.0, tol, 25, False)
        margin = min(min(x[-2] / tol, tol / max(x[-1], 1e-300)) for x in h)
        print(tol, margin, [len(x) for x in h])
s = soil_column_equaldof(12, mat=J2_STEEL, distort=0.1)
s.loads = np.array([[1 + 2 * 12, 22.0, -3.0], [1 + 2 * 6, 10.0, 0.0]])
run(s, 8)
run(frame2d_diaphragm_equaldof(2, 3, 2, lateral=22.0, gravity=-40.0), 5)
