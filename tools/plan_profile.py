"""Phases of the cold plan build (ordering + symbolic analysis + schedule) of the dense 100 k world, host only.
    python tools/plan_profile.py            (ASAM_PLAN_THREADS=1: serial symbolic loops)"""
import os, sys, time, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from aprilsam_b200 import datasets
from support.hostplan import HostPlan, lib
L=lib()
d=datasets.manhattan_dense(100000, seed=1)
E=d.n_edges
ftype=np.ones(E+1,np.int32); ftype[0]=2
fa=np.concatenate([[0],d.ea]).astype(np.int32); fb=np.concatenate([[-1],d.eb]).astype(np.int32)
prof=(C.c_double*8)()
L.asam_dbg_build_profile_get.argtypes=[C.POINTER(C.c_double), C.c_int]
best=None
for it in range(4):
    L.asam_dbg_build_profile_get(prof,1)
    p=HostPlan(); t=time.time(); p.build(d.n_nodes,ftype,fa,fb); dt=(time.time()-t)*1e3
    L.asam_dbg_build_profile_get(prof,0)
    v=[prof[i] for i in range(8)]
    print("build %.1f ms: slots+adj %.1f ordering %.1f block-symbolic %.1f postorder+sn %.1f rows+rel %.1f gather %.1f seg+schedule %.1f upload %.1f"%(dt,*v))
