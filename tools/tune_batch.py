#!/usr/bin/env python
"""One configuration of the tuning knobs (environment: ASAM_SOLO_MAX_M, ASAM_SOLO_PB, ASAM_TEAM_ROOM, ...)
on the dense synthetic world: kernel times of warm batch solves + the solution checked against a stored
one (run under gpurun, one process per configuration; see tools/tune_sweep.sh).

    python tools/tune_batch.py --poses 100000 --tag base --save /tmp/x.npy
    ASAM_SOLO_MAX_M=330 python tools/tune_batch.py --poses 100000 --tag solo330 --check /tmp/x.npy
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aprilsam_b200 import capi, datasets  # noqa: E402
from aprilsam_b200 import harness as H  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=100000)
    ap.add_argument("--workload", default="dense", choices=["dense", "sparse", "m3500"])
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--tag", default="")
    ap.add_argument("--save", default="")
    ap.add_argument("--check", default="")
    args = ap.parse_args()
    L = capi.lib()
    if args.workload == "m3500":
        d = H.PoseGraphData.load(os.path.join(ROOT, "tests", "golden", "m3500.npz"))
    elif args.workload == "sparse":
        d = datasets.manhattan_sparse(args.poses, seed=1)
    else:
        d = datasets.manhattan_dense(args.poses, seed=1)
    with H.Harness("b200") as h:
        h.load_full(d)
        h.batch()
        st = h.states()
        dev = L.asam_dbg_dev_of_graph(h.graph_ptr())
        L.asam_set_timing(dev, 1)
        km, e2e = [], []
        prof = (C.c_double * 24)()
        L.asam_dbg_profile(prof, 1)
        for _ in range(args.iters):
            h.set_states(d.init)
            e2e.append(h.batch())
            km.append(capi.kernel_ms(dev))
        km = np.median(np.array(km), axis=0)
        L.asam_dbg_profile(prof, 1)
        host = " host ms/call: gather %.3f plan-check %.3f enqueue+verify %.3f wait+D2H %.3f tree+update %.3f" % tuple(
            prof[i] / args.iters for i in (11, 12, 13, 14, 16))
    err = ""
    if args.save:
        np.save(args.save, st)
    if args.check and os.path.exists(args.check):
        ref = np.load(args.check)
        dd = st - ref
        dd[:, 2] = (dd[:, 2] + np.pi) % (2 * np.pi) - np.pi
        err = f" max_rel_vs_base {float(np.max(np.abs(dd) / np.maximum(1.0, np.abs(ref)))):.2e}"
    knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("ASAM_"))
    print(f"TUNE {args.tag or 'cfg'} [{knobs}] lin {km[0]:.3f} factor {km[1]:.3f} backsolve {km[2]:.3f} ms; e2e median {np.median(e2e):.3f} ms{err};{host}", flush=True)


if __name__ == "__main__":
    main()
