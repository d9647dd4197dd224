"""ctypes wrapper of oracle/aprilsam_oracle.c (TEST INFRASTRUCTURE: the plain-C restatement of the
reference's batch step + chi2).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
may import this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "liboracle_port.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.oracle_chi2.restype = C.c_double
        L.oracle_chi2.argtypes = [C.c_int, _dp, C.c_int, _ip, _ip, _ip, _dp, _dp]
        L.oracle_batch_step.restype = C.c_int
        L.oracle_batch_step.argtypes = [C.c_int, _dp, C.c_int, _ip, _ip, _ip, _dp, _dp, C.c_double, _dp]
        _lib = L
    return _lib


def available() -> bool:
    return os.path.exists(_PATH)


def _factors(d):
    """Demo protocol: xytpos prior on node 0 (W = diag(1e4,1e4,1e3), z = 0) + every EDGE2 as xyt."""
    E = d.n_edges
    ftype = np.ascontiguousarray(np.r_[2, np.ones(E)], dtype=np.int32)
    fa = np.ascontiguousarray(np.r_[0, d.ea], dtype=np.int32)
    fb = np.ascontiguousarray(np.r_[-1, d.eb], dtype=np.int32)
    fz = np.ascontiguousarray(np.vstack([[0, 0, 0], d.ez]), dtype=np.float64)
    fW = np.ascontiguousarray(np.vstack([[1e4, 0, 0, 0, 1e4, 0, 0, 0, 1e3], d.eW]), dtype=np.float64)
    return ftype, fa, fb, fz, fW


def chi2(d, states) -> float:
    ftype, fa, fb, fz, fW = _factors(d)
    st = np.ascontiguousarray(states, dtype=np.float64)
    return lib().oracle_chi2(d.n_nodes, st.ctypes.data_as(_dp), len(ftype), ftype.ctypes.data_as(_ip),
                             fa.ctypes.data_as(_ip), fb.ctypes.data_as(_ip), fz.ctypes.data_as(_dp),
                             fW.ctypes.data_as(_dp))


def batch_step(d, states, lam: float = 1e-4) -> np.ndarray:
    """One april_graph_cholesky-equivalent step from `states`; returns the new states."""
    ftype, fa, fb, fz, fW = _factors(d)
    st = np.array(states, dtype=np.float64, order="C", copy=True)
    rc = lib().oracle_batch_step(d.n_nodes, st.ctypes.data_as(_dp), len(ftype), ftype.ctypes.data_as(_ip),
                                 fa.ctypes.data_as(_ip), fb.ctypes.data_as(_ip), fz.ctypes.data_as(_dp),
                                 fW.ctypes.data_as(_dp), lam, None)
    if rc:
        raise np.linalg.LinAlgError("oracle port: matrix not positive definite")
    return st
