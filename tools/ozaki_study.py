#!/usr/bin/env python
"""Numerical study for the tensor-pipe question (host only, numpy): what would an Ozaki-split
trailing update cost on this path?  The multifrontal factorisation of the M3500 batch step is emulated (tests/support/emul.py,
the code the CPU tests use to check the uploaded plan) with panels factored in double precision and every trailing product
P P' computed from s slices of 7-bit integers per operand entry (int8 tensor-core operands, exact int32 accumulation; the
slice pairs (t, u) with t + u < s are kept), and the node states are compared with the all-double emulation.

    python tools/ozaki_study.py [panel width, default 48] [poses of a synthetic dense world instead of M3500]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aprilsam_b200 import harness as H  # noqa: E402
from support import emul  # noqa: E402
from support.hostplan import HostPlan  # noqa: E402

BITS = 7


def ozaki_syrk(P, s):
    """P P' from s slices of BITS-bit integers per entry (row-wise power-of-two scaling)."""
    n, k = P.shape
    amax = np.abs(P).max(axis=1)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)
    Y = P * np.exp2(-e)[:, None]  # |Y| < 1
    slices, r = [], Y.copy()
    for _ in range(s):
        r = r * float(1 << BITS)
        q = np.trunc(r)
        r = r - q
        slices.append(q)  # integers in (-2^BITS, 2^BITS)
    acc = np.zeros((n, n))
    for t in range(s):
        for u in range(s - t):
            acc += (slices[t] @ slices[u].T) * 2.0 ** (-BITS * (t + u + 2))  # exact: |sum| < k * 2^14 < 2^31
    return acc * np.exp2(e)[:, None] * np.exp2(e)[None, :]


def solve(d, plan, panel, syrk):
    n = d.n_nodes
    ftype, fa, fb, fz, fW = arrays(d)
    info = plan.info()
    Hs = emul.Hessian(n, info["n_slots"])
    Hs.reset(n, 1e-4)
    node2q = plan.array("node2q")
    Hs.linearize(range(len(ftype)), ftype, fa, fb, fz, fW, d.init, d.init, node2q, plan.array("fslot"))
    fr = emul.Fronts()
    fr.ensure(n)
    desc, ipool = plan.descs(), plan.array("ipool")
    leaf = plan.array("leaf_tasks")
    if len(leaf):
        emul.factor(fr, Hs, desc, ipool, plan.array("q2node"), leaf, None, panel=panel, syrk=syrk)
    emul.factor(fr, Hs, desc, ipool, plan.array("q2node"), plan.array("tasks"), plan.array("nwait"), prior=leaf, panel=panel, syrk=syrk)
    emul.backsolve(fr, desc, ipool, plan.array("btasks"))
    x = np.stack([fr.x[3 * node2q[i]:3 * node2q[i] + 3] for i in range(n)])
    st = d.init + x
    st[:, 2] = emul.mod2pi(st[:, 2])
    return st


def arrays(d):
    E = d.n_edges
    ftype = np.ones(E + 1, np.int32); ftype[0] = 2
    fa = np.concatenate([[0], d.ea]).astype(np.int32); fb = np.concatenate([[-1], d.eb]).astype(np.int32)
    fz = np.vstack([[0, 0, 0], d.ez]); fW = np.vstack([[1e4, 0, 0, 0, 1e4, 0, 0, 0, 1e3], d.eW])
    return ftype, fa, fb, fz, fW


def main():
    panel = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    if len(sys.argv) > 2:  # a dense synthetic world of that many poses instead (team-sized fronts from ~5000 on)
        from aprilsam_b200 import datasets
        d = datasets.manhattan_dense(int(sys.argv[2]), seed=1)
        name = f"synthetic dense Manhattan world, {sys.argv[2]} poses"
    else:
        d = H.PoseGraphData.load(os.path.join(ROOT, "tests", "golden", "m3500.npz"))
        name = "M3500"
    ftype, fa, fb, _, _ = arrays(d)
    plan = HostPlan().build(d.n_nodes, ftype, fa, fb)
    ref = solve(d, plan, panel, None)
    scale = max(1.0, np.abs(ref).max())
    print(f"{name}: batch step, panels of {panel} columns in double, trailing products from int{BITS + 1} slices (numpy emulation)")
    print("slices  int8 products per update   max |state - double| / max|state|   within 1e-6?")
    for s in range(2, 10):
        try:
            st = solve(d, plan, panel, lambda P, s=s: ozaki_syrk(P, s))
            err = np.abs(st - ref)
            err[:, 2] = np.abs((err[:, 2] + np.pi) % (2 * np.pi) - np.pi)
            rel = err.max() / scale
            print(f"{s:6d}  {s * (s + 1) // 2:24d}   {rel:34.3e}   {'yes' if rel < 1e-6 else 'no'}")
        except np.linalg.LinAlgError as ex:
            print(f"{s:6d}  {s * (s + 1) // 2:24d}   factorisation breaks down ({ex})")


if __name__ == "__main__":
    main()
