"""Structure of the elimination tree the (reference-equivalent) ordering gives a synthetic world, CPU only:
fill, flops, tree height and the dependent chain of k_factor (team fronts / 48-column panel steps on the
heaviest root path, with each front's share of the work: what a multi-GPU cut can and cannot split).
python tools/order_stats.py [N] [dense|sparse]     (CHAIN=1: list the fronts on the chain)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from aprilsam_b200 import datasets
from support.hostplan import HostPlan, lib

def factor_arrays(d):
    E = d.n_edges
    ftype = np.ones(E + 1, np.int32); ftype[0] = 2
    fa = np.concatenate([[0], d.ea]).astype(np.int32); fb = np.concatenate([[-1], d.eb]).astype(np.int32)
    return ftype, fa, fb

def stats(p, label, dt):
    D = p.descs(); info = p.info()
    m = 3 * D["mb"].astype(np.int64); c = 3 * D["cb"].astype(np.int64); par = D["parent"]
    fits = ((m + 2) // 2 * 2) * m + ((m + 2) // 2 * 2 + 1) // 2 + 2 <= 25600
    # per-front latency model (us): team fronts 30/panel of 48 cols, smem fronts 4 + 1.2/panel of 12 cols
    lat = np.where(fits, 4.0 + 1.2 * np.ceil(c / 12.0) + 0.02 * m, 10.0 + 30.0 * np.ceil(c / 48.0))
    nsn = len(m); path = np.zeros(nsn); pan = np.zeros(nsn)
    for s in range(nsn):  # children have smaller ids
        path[s] += lat[s]; pan[s] += 0 if fits[s] else np.ceil(c[s] / 48.0)
        if par[s] >= 0:
            if path[s] > path[par[s]]: path[par[s]] = path[s]; pan[par[s]] = pan[s]
    roots = np.where(par < 0)[0]
    r = roots[np.argmax(path[roots])]
    work = (c * m * m).astype(float)
    print(f"{label:10s} build {dt*1e3:7.1f} ms  nsn {nsn:6d} levels {info['n_levels']:3d} max_m {info['max_m']:5d} "
          f"nnzL {9*info['nnz_l_blocks']/1e6:7.2f} M  flops {info['flops']/1e9:7.3f} G  arena {info['arena_n']*8/1e6:7.1f} MB  "
          f"team fronts {int((~fits).sum()):4d}  chain ~{path[r]/1e3:6.2f} ms ({int(pan[r])} team panels)  "
          f"team work {work[~fits].sum()/work.sum()*100:4.1f}%")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
kind = sys.argv[2] if len(sys.argv) > 2 else "dense"
d = (datasets.manhattan_dense if kind == "dense" else datasets.manhattan_sparse)(N, seed=1)
ftype, fa, fb = factor_arrays(d)
p = HostPlan()
t = time.time(); p.build(d.n_nodes, ftype, fa, fb); dt = time.time() - t
stats(p, "ref-MD", dt)

def chain(p, label):
    D = p.descs()
    m = 3 * D["mb"].astype(np.int64); c = 3 * D["cb"].astype(np.int64); par = D["parent"]
    fits = ((m + 2) // 2 * 2) * m + ((m + 2) // 2 * 2 + 1) // 2 + 2 <= 25600
    lat = np.where(fits, 4.0 + 1.2 * np.ceil(c / 12.0) + 0.02 * m, 10.0 + 30.0 * np.ceil(c / 48.0))
    nsn = len(m); path = lat.copy(); via = -np.ones(nsn, int)
    for s in range(nsn):
        if par[s] >= 0 and path[s] + lat[par[s]] > path[par[s]]:
            path[par[s]] = path[s] + lat[par[s]]; via[par[s]] = s
    # subtree work
    sub = (c * m * m).astype(float)
    for s in range(nsn):
        if par[s] >= 0: sub[par[s]] += sub[s]
    r = int(np.argmax(np.where(par < 0, path, -1)))
    print(label, "critical chain root->leaf (m, c, subtree work share, #children):")
    kids = np.bincount(par[par >= 0], minlength=nsn)
    s = r; tot = sub[r]; k = 0
    while s >= 0 and k < 40:
        if not fits[s]:
            print(f"   sn {s:6d} m {m[s]:5d} c {c[s]:5d} sub {sub[s]/tot*100:5.1f}% kids {kids[s]}")
            k += 1
        s = via[s]
if os.environ.get("CHAIN"):
    chain(p, "ref-MD")
