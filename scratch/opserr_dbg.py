import sys, faulthandler; faulthandler.enable()
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import ctypes, numpy as np
from modelspec import *
L=ctypes.CDLL(REF_SO)
L.ref_model_new.restype=ctypes.c_void_p
h=ctypes.c_void_p(L.ref_model_new(3,3))
x=np.zeros(3)
print(L.ref_add_node(h,1,x.ctypes.data_as(ctypes.c_void_p)))
print('fix1', L.ref_fix(h,1,0)); sys.stdout.flush()
print('fix2', L.ref_fix(h,1,0)); sys.stdout.flush()
