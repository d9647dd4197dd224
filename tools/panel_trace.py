#!/usr/bin/env python
"""Where does a 48-column panel step of k_factor's team path spend its time?  (run under gpurun)

Loads the dense synthetic world, runs one batch solve with asam_set_panel_trace on the widest team
fronts of the tree's critical path and prints, per front, the median duration of the phases of a panel
step for crew worker 0 (diagonal block), crew worker 1 (first row chunk) and a trailing-update worker.

    python tools/panel_trace.py [--poses 100000] [--fronts 3]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aprilsam_b200 import capi, datasets  # noqa: E402
from aprilsam_b200 import harness as H  # noqa: E402
from support.hostplan import HostPlan  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=100000)
    ap.add_argument("--fronts", type=int, default=3)
    ap.add_argument("--m3500", action="store_true", help="the M3500 fixture instead of the synthetic world")
    ap.add_argument("--dump-trace", default="", help="npz file: per-task stamps of one k_factor / k_backsolve launch + the plan")
    args = ap.parse_args()
    L = capi.lib()
    L.asam_set_panel_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.asam_download_panel_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
    d = (H.PoseGraphData.load(os.path.join(ROOT, "tests", "golden", "m3500.npz")) if args.m3500
         else datasets.manhattan_dense(args.poses, seed=1))
    # the plan (host only) to pick the fronts: widest supernodes of the team path
    E = d.n_edges
    ftype = np.ones(E + 1, np.int32); ftype[0] = 2
    fa = np.concatenate([[0], d.ea]).astype(np.int32); fb = np.concatenate([[-1], d.eb]).astype(np.int32)
    D = HostPlan().build(d.n_nodes, ftype, fa, fb).descs()
    m, c = 3 * D["mb"].astype(np.int64), 3 * D["cb"].astype(np.int64)
    team = ((m + 2) // 2 * 2) * m + ((m + 2) // 2 * 2 + 1) // 2 + 2 > 25600
    order = np.argsort(-(c * team))[:args.fronts]
    with H.Harness("b200") as h:
        h.load_full(d)
        h.batch()
        dev = L.asam_dbg_dev_of_graph(h.graph_ptr())
        L.asam_set_timing(dev, 1)
        for s in order:
            if not team[s]:
                continue
            npan = int((c[s] + 47) // 48)
            capi.check(L.asam_set_panel_trace(dev, int(s), npan), "asam_set_panel_trace")
            h.set_states(d.init)
            h.batch()
            km = capi.kernel_ms(dev)
            buf = np.zeros((npan, 8, 8), dtype=np.uint64)
            capi.check(L.asam_download_panel_trace(dev, buf.ctypes.data_as(C.POINTER(C.c_uint64)), npan), "download")
            t = buf.astype(np.int64)
            print(f"front sn {s}: m {m[s]} c {c[s]} ({npan} panels), team {t[0, 0, 6]} CTAs; k_factor {km[1]:.3f} ms", flush=True)
            w0 = t[:, 0, :]
            step = w0[:, 4] - w0[:, 0]
            print(f"  panel step (worker 0, start -> past barrier): median {np.median(step) / 1e3:.1f} us, "
                  f"first {step[0] / 1e3:.1f}, last {step[-1] / 1e3:.1f}, sum {step.sum() / 1e3:.0f} us; "
                  f"front total (first start -> last barrier) {(w0[-1, 4] - w0[0, 0]) / 1e3:.0f} us")
            for w, name in ((0, "worker 0 (diag block)"), (1, "worker 1 (row chunk)"), (7, "worker 7")):
                tw = t[:, w, :]
                ok = tw[:, 0] > 0
                if not ok.any():
                    continue
                tw = tw[ok]
                ph = np.stack([tw[:, 1] - tw[:, 0], tw[:, 2] - tw[:, 1], tw[:, 3] - tw[:, 2], tw[:, 4] - tw[:, 3]], 1) / 1e3
                print(f"  {name:24s} crew-tile {np.median(ph[:, 0]):6.1f}  factor/solve {np.median(ph[:, 1]):6.1f}  "
                      f"trailing {np.median(ph[:, 2]):6.1f}  barrier-wait {np.median(ph[:, 3]):6.1f}   (us, medians over {len(tw)} panels; "
                      f"crew size first/last {tw[0, 7]}/{tw[-1, 7]})")
            # per panel table for the first front
            if s == order[0]:
                print("  panel:  step_us  w0[tile,factor+publish,trailing,barrier]  w1[tile,wait+solve,trailing,barrier]  rows_left")
                for k in range(npan):
                    a, b = t[k, 0], t[k, 1]
                    fa_ = [(a[1] - a[0]) / 1e3, (a[2] - a[1]) / 1e3, (a[3] - a[2]) / 1e3, (a[4] - a[3]) / 1e3]
                    fb_ = [(b[1] - b[0]) / 1e3, (b[2] - b[1]) / 1e3, (b[3] - b[2]) / 1e3, (b[4] - b[3]) / 1e3] if b[0] > 0 else [0] * 4
                    print(f"   {k:3d}  {(a[4] - a[0]) / 1e3:7.1f}   " + " ".join(f"{x:6.1f}" for x in fa_) + "    " +
                          " ".join(f"{x:6.1f}" for x in fb_) + f"    {m[s] - 48 * (k + 1)}")
        capi.check(L.asam_set_panel_trace(dev, -1, 0), "off")
        if args.dump_trace:
            L.asam_set_trace.argtypes = [C.c_void_p, C.c_int]
            L.asam_download_trace.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_int]
            P = HostPlan().build(d.n_nodes, ftype, fa, fb)
            tasks, nwait, btasks = P.array("tasks"), P.array("nwait"), P.array("btasks")
            L.asam_set_trace(dev, 1)
            h.set_states(d.init)
            h.batch()
            km = capi.kernel_ms(dev)
            tf = np.zeros((len(tasks), 8), dtype=np.uint64)
            tb = np.zeros((len(btasks), 8), dtype=np.uint64)
            L.asam_download_trace(dev, 0, tf.ctypes.data_as(C.POINTER(C.c_uint64)), len(tasks))
            L.asam_download_trace(dev, 1, tb.ctypes.data_as(C.POINTER(C.c_uint64)), len(btasks))
            L.asam_set_trace(dev, 0)
            np.savez_compressed(args.dump_trace, tf=tf, tb=tb, tasks=tasks, nwait=nwait, btasks=btasks, leaf=P.array("leaf_tasks"),
                                kernel_ms=np.array(km), **{k: v for k, v in D.items()})
            print(f"trace of {len(tasks)} factor tasks / {len(btasks)} back-solve tasks -> {args.dump_trace}; kernels ms {km}")


if __name__ == "__main__":
    main()
