#!/usr/bin/env python
"""Bisect edge-case parity on the GPU box: variants of a small graph, b200 vs reference."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aprilsam_b200 import harness as H


def rel_state_err(a, b):
    d = a - b
    d[:, 2] = (d[:, 2] + np.pi) % (2 * np.pi) - np.pi
    return float(np.max(np.abs(d) / np.maximum(1.0, np.abs(b))))


def run(name, extra, asym, prior2, tik, inc):
    rng = np.random.default_rng(5)
    n0, n1 = 30, 38
    truth = np.cumsum(np.c_[np.ones(n1), 0.3 * rng.standard_normal(n1), 0.2 * rng.standard_normal(n1)], axis=0)

    def rel(a, b):
        c, s = np.cos(truth[a, 2]), np.sin(truth[a, 2])
        d = truth[b] - truth[a]
        return np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], d[2]]) + 0.01 * rng.standard_normal(3)

    def full_W():
        M = rng.standard_normal((3, 3))
        W = 30.0 * (M @ M.T + 0.5 * np.eye(3))
        if asym:
            W[0, 1] += 0.05
            W[1, 2] -= 0.03
        return W

    edges = [(i, i + 1) for i in range(n1 - 1)]
    recs = {e: (rel(*e), full_W()) for e in sorted(set(edges + extra))}
    init = truth + 0.05 * rng.standard_normal(truth.shape)
    Wp2 = full_W()

    def drive(h):
        out = []
        if tik:
            h.set_tikhanov(3e-3)
        for k in range(n0):
            h.add_node(init[k])
        h.add_xytpos(0, truth[0], np.diag([1e4, 1e4, 1e3]))
        if prior2:
            h.add_xytpos(17, truth[17] + 0.01, Wp2)
        for (a, b) in edges + extra:
            if max(a, b) < n0:
                h.add_xyt(a, b, *recs[(a, b)])
        for it in range(3):
            h.batch()
            out.append((h.states(), h.chi2(), 0))
        for k in range(n0, n1 if inc else n0):
            h.add_node(init[k])
            more = []
            if inc in (2, 4):
                more.append((k - 4, k))
            if inc in (3, 4):
                more.append((k, k - 9))
            for (a, b) in edges + more:
                if max(a, b) == k:
                    if (a, b) not in recs:
                        recs[(a, b)] = (rel(a, b), full_W())
                    h.add_xyt(a, b, *recs[(a, b)])
            h.inc()
            out.append((h.states(), h.chi2(), h.info()["naffected"]))
        return out

    with H.Harness("reference", nthreshold=10**9) as b:
        rb = drive(b)
    if "--ref-only" in sys.argv:
        print(f"{name:40s} reference ok, final chi2 {rb[-1][1]:.6g}", flush=True)
        return
    with H.Harness("b200", nthreshold=10**9) as a:
        ra = drive(a)
    errs = [rel_state_err(sa, sb) for (sa, _, _), (sb, _, _) in zip(ra, rb)]
    print(f"{name:40s} max err {max(errs):.3e}  per step {['%.1e' % e for e in errs]}", flush=True)


VARIANTS = {}
def variant(name, *args):
    VARIANTS[name] = args

base = [(3, 11), (7, 20), (2, 25)]
variant("forward closures, sym W", base, False, False, False, False)
variant("+ asym W + tikhanov (asym alone: reference goes indefinite and crashes)", base, True, False, True, False)
variant("+ reversed pair (20,7)", [(3, 11), (20, 7), (2, 25)], False, False, False, False)
variant("+ adjacent reversed (14,13)", base + [(14, 13)], False, False, False, False)
variant("+ duplicate (3,11) x2", base + [(3, 11)], False, False, False, False)
variant("+ both directions (3,11),(11,3)", base + [(11, 3)], False, False, False, False)
variant("+ second prior", base, False, True, False, False)
variant("+ tikhanov 3e-3", base, False, False, True, False)
variant("+ incremental appends", base, False, False, False, True)
variant("inc + (k-4,k)", base, False, False, False, 2)
variant("inc + (k,k-9) new pose first", base, False, False, False, 3)
variant("inc + both closures", base, False, False, False, 4)
variant("inc + both closures, asym+tik", base, True, False, True, 4)
variant("everything but asym", [(3, 11), (11, 3), (3, 11), (20, 7), (25, 2), (14, 13), (29, 0), (28, 9)], False, True, True, True)
variant("everything", [(3, 11), (11, 3), (3, 11), (20, 7), (25, 2), (14, 13), (29, 0), (28, 9)], True, True, True, True)


if __name__ == "__main__":
    import subprocess
    only_ref = "--ref-only" in sys.argv
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        name = sys.argv[2]
        if only_ref:
            H_b200 = H.Harness
            class _Fake:
                pass
        run(name, *VARIANTS[name])
        sys.exit(0)
    for name in VARIANTS:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name] + (["--ref-only"] if only_ref else []),
                           capture_output=True, text=True)
        if r.returncode != 0:
            print(f"{name:40s} CRASH rc={r.returncode}: {(r.stderr.strip().splitlines() or ['?'])[-1][:150]}", flush=True)
        else:
            print(r.stdout.strip(), flush=True)
