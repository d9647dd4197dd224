#!/usr/bin/env python
"""Lock-step check of the incremental path at scale (GPU box): bulk-load the first S poses of the
sparse synthetic Manhattan graph, one batch solve, then K incremental steps on both arms
(aprilsam_b200 vs oracle/_ref), comparing naffected / start_over / states every `chunk` steps.

    python tools/inc_scale_check.py --poses 100000 --start 99000 --steps 60 --chunk 5
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aprilsam_b200 import datasets  # noqa: E402
from aprilsam_b200 import harness as H  # noqa: E402


def rel_err(a, b):
    d = a - b
    d[:, 2] = (d[:, 2] + np.pi) % (2 * np.pi) - np.pi
    return float(np.abs(d).max() / max(1.0, np.abs(b).max()))


def exact_check(a, b, d):
    """x_exact = A^-1 B from the Hessian in HBM (linearised at the l_points); which arm's
    (state - l_point) on the poses touched by the last steps is closer?"""
    import ctypes as C
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from aprilsam_b200 import capi
    L = capi.lib()
    dev = L.asam_dbg_dev_of_graph(a.graph_ptr())
    pinfo = capi.plan_info(L.asam_dbg_plan_of_param(a.param_ptr()))
    N, S = pinfo["N"], pinfo["n_slots"]
    Ad = np.zeros((N, 3, 3)); Ao = np.zeros((S, 3, 3)); B = np.zeros((N, 3))
    dp = C.POINTER(C.c_double)
    capi.check(L.asam_debug_read_hessian(dev, N, S, Ad.ctypes.data_as(dp), Ao.ctypes.data_as(dp), B.ctypes.data_as(dp)))
    # slot numbering = first appearance of the (lo, hi) pair in factor order (prior first, then edges as added)
    db, estart = d.bucketed()
    n_bulk = None
    ea = np.concatenate([d.head(args_start).ea, db.ea[estart[args_start]:estart[N]]])
    eb = np.concatenate([d.head(args_start).eb, db.eb[estart[args_start]:estart[N]]])
    lo = np.minimum(ea, eb).astype(np.int64); hi = np.maximum(ea, eb).astype(np.int64)
    key = lo * N + hi
    _, first_idx = np.unique(key, return_index=True)
    order = np.sort(first_idx)
    plo, phi = lo[order], hi[order]
    assert len(plo) == S, (len(plo), S)
    rows, cols, vals = [], [], []
    iu = np.triu_indices(3)
    D = np.zeros_like(Ad)
    D[:, iu[0], iu[1]] = Ad[:, iu[0], iu[1]]
    D = D + np.transpose(np.triu(D, 1), (0, 2, 1))
    base = 3 * np.arange(N)
    for p in range(3):
        for q in range(3):
            rows.append(base + p); cols.append(base + q); vals.append(D[:, p, q])
            rows.append(3 * plo + p); cols.append(3 * phi + q); vals.append(Ao[:, p, q])
            rows.append(3 * phi + q); cols.append(3 * plo + p); vals.append(Ao[:, p, q])
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(3 * N, 3 * N))
    t0 = time.time()
    lu = spl.splu(A, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    x = lu.solve(B.reshape(-1))
    for _ in range(2):  # iterative refinement in double
        x += lu.solve(B.reshape(-1) - A @ x)
    print(f"exact solve: {time.time() - t0:.1f} s, residual {np.abs(A @ x - B.reshape(-1)).max():.3e}", flush=True)
    x = x.reshape(N, 3)
    tail = slice(max(0, N - 8), N)
    xa = a.states() - a.l_points()
    xa[:, 2] = (xa[:, 2] + np.pi) % (2 * np.pi) - np.pi
    xe = x.copy(); xe[:, 2] = (xe[:, 2] + np.pi) % (2 * np.pi) - np.pi
    print("only poses marked by the LAST step hold its solution (naffected <= 5 prunes the rest); per pose:", flush=True)
    xb = None
    if b:
        xb = b.states() - b.l_points()
        xb[:, 2] = (xb[:, 2] + np.pi) % (2 * np.pi) - np.pi
    for i in range(max(0, N - 4), N):
        line = f"  pose {i}: |x| {np.abs(xe[i]).max():.3e}  |b200-exact| {np.abs(xa[i] - xe[i]).max():.3e}"
        if xb is not None:
            line += f"  |ref-exact| {np.abs(xb[i] - xe[i]).max():.3e}  |b200-ref| {np.abs(xa[i] - xb[i]).max():.3e}"
        print(line, flush=True)


args_start = 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=30000)
    ap.add_argument("--start", type=int, default=29500)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--chunk", type=int, default=20)
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--dead-reckoned", action="store_true", help="start the bulk load at the dead-reckoned VERTEX2 estimate")
    ap.add_argument("--exact", action="store_true",
                    help="after the steps: solve the Hessian held in HBM with scipy (SuperLU) and compare both arms")
    args = ap.parse_args()
    global args_start
    args_start = args.start
    d = datasets.manhattan_dense(args.poses, seed=1) if args.dense else datasets.manhattan_sparse(args.poses, seed=1)
    have_ref = H.available("reference")
    a = H.Harness("b200")
    b = H.Harness("reference") if have_ref else None
    for h in (a, b):
        if h is None:
            continue
        h.replay_begin(d)
        t0 = time.time()
        sub = d.head(args.start)
        h.load_full(sub)
        if sub.truth is not None and not args.dead_reckoned:
            h.set_states(sub.truth)
        h.batch()
        h.batch()
        print(f"[{h.impl}] bulk {args.start} poses + batch: {time.time() - t0:.2f} s", flush=True)
    if b:
        print(f"after batch: rel state err {rel_err(a.states(), b.states()):.3e}", flush=True)
    worst = 0.0
    for k in range(args.start + args.chunk, args.start + args.steps + 1, args.chunk):
        _, msa, ia = a.replay_to(k, want_chi2=False)
        line = f"step {k}: b200 {msa.mean():.3f} ms/step naff max {ia[:, 0].max()}"
        if b:
            _, msb, ib = b.replay_to(k, want_chi2=False)
            e = rel_err(a.states(), b.states())
            worst = max(worst, e)
            same = np.array_equal(ia[:, 0], ib[:, 0]) and np.array_equal(ia[:, 1], ib[:, 1])
            line += f" | ref {msb.mean():.3f} ms/step | naff/start_over equal {same} | rel state err {e:.3e}"
        print(line, flush=True)
    print(f"worst rel state err {worst:.3e}", flush=True)
    if args.exact:
        exact_check(a, b, d)
    a.close()
    if b:
        b.close()
    return 0 if worst < 1e-6 else 1


if __name__ == "__main__":
    sys.exit(main())
