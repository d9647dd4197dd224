#!/usr/bin/env python
"""Host-phase profile of the incremental path (run under gpurun): M3500 demo replay, per-phase host time
per april_graph_cholesky_inc call from the library's own lap timers, split by step size.

    python tools/step_profile.py [--steps 3500]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aprilsam_b200 import capi  # noqa: E402
from aprilsam_b200 import harness as H  # noqa: E402

NAMES = ["pre: sync factors, grow tree, mark", "plan_append (+queue uploads)", "linearize+factor record", "tree append",
         "backsolve list + launch + wait + x", "status", "apply_solution"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3500)
    args = ap.parse_args()
    L = capi.lib()
    d = H.PoseGraphData.load(os.path.join(ROOT, "tests", "golden", "m3500.npz"))
    prof = (C.c_double * 24)()
    with H.Harness("b200") as h:
        h.replay_begin(d)
        h.replay_to(1, want_chi2=False)
        L.asam_dbg_profile(prof, 1)
        rows = []
        for k in range(2, min(args.steps, d.n_nodes) + 1):
            _, ms, info = h.replay_to(k, want_chi2=False)
            L.asam_dbg_profile(prof, 1)
            rows.append([ms[0] * 1e3, info[0, 0]] + [prof[i] * 1e3 for i in range(7)] + [prof[18], prof[10], prof[19], prof[20], prof[21]])
        dev = L.asam_dbg_dev_of_graph(h.graph_ptr())
        L.asam_small_steps.argtypes = [C.c_void_p]
        L.asam_small_steps.restype = C.c_int64
        ns = L.asam_small_steps(dev)
        print("fused small steps:", ns, "of", len(rows))
        L.asam_small_step_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        sp = (C.c_double * 7)()
        L.asam_small_step_profile(dev, sp, 0)
        if ns:
            nm = ["fetch uploads + scatter", "linearize", "factor", "back-solve", "write results", "host: launch call", "host: wait on flag"]
            print("k_step phases, mean us per fused step: " + ", ".join(f"{n} {sp[i] / ns:.2f}" for i, n in enumerate(nm)))
    r = np.array(rows)
    for name, sel in (("naffected <= 5, fused", (r[:, 1] <= 5) & (r[:, 1] > 0) & (r[:, 9] > 0)),
                      ("naffected <= 5, general path", (r[:, 1] <= 5) & (r[:, 1] > 0) & (r[:, 9] == 0)),
                      ("naffected 6..50", (r[:, 1] > 5) & (r[:, 1] <= 50)), ("naffected > 50", r[:, 1] > 50)):
        x = r[sel]
        if not len(x):
            continue
        print(f"{name}: {len(x)} steps, call median {np.median(x[:, 0]):.1f} us, mean {x[:, 0].mean():.1f} us"
              + (f"; fronts re-factored mean {x[:, 11].mean():.2f}, supernodes back-solved mean {x[:, 12].mean():.2f}, x doubles mean {x[:, 13].mean():.1f}"
                 if x[:, 9].sum() > 0 else ""))
        for i, nm in enumerate(NAMES):
            print(f"    {nm:40s} median {np.median(x[:, 2 + i]):7.2f}  mean {x[:, 2 + i].mean():7.2f} us")


if __name__ == "__main__":
    main()
