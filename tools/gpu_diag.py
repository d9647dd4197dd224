#!/usr/bin/env python
"""GPU bring-up diagnostics (run under gpurun): kernel-level comparison of what is in HBM
against the numpy emulation (tests/support/emul.py) and state-level comparison against the
reference, with enough detail in the log to locate a defect without another round trip.

    python tools/gpu_diag.py [--sizes 2,10,100,3500] [--replay 300] [--time]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from aprilsam_b200 import harness as H  # noqa: E402
from support import emul  # noqa: E402
from support.hostplan import HostPlan, lib as hostlib  # noqa: E402

_dp = C.POINTER(C.c_double)
DUMP = {}


def log(*a):
    print(*a, flush=True)


def dev_api():
    L = hostlib()
    L.asam_dbg_dev_of_graph.argtypes = [C.c_void_p]
    L.asam_dbg_dev_of_graph.restype = C.c_void_p
    L.asam_dbg_plan_of_param.argtypes = [C.c_void_p]
    L.asam_dbg_plan_of_param.restype = C.c_void_p
    L.asam_debug_read_hessian.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
    L.asam_download_x.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    L.asam_download_y.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    L.asam_debug_read_front.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _dp]
    L.asam_last_error.restype = C.c_char_p
    L.asam_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.asam_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.asam_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.asam_set_trace.argtypes = [C.c_void_p, C.c_int]
    L.asam_download_trace.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_int]
    return L


def trace_report(L, dev, ntasks, label):
    for which, name in ((0, "k_factor"), (1, "k_backsolve")):
        tr = np.zeros((ntasks, 8), dtype=np.uint64)
        L.asam_download_trace(dev, which, tr.ctypes.data_as(C.POINTER(C.c_uint64)), ntasks)
        tr = tr[tr[:, 0] > 0]
        if len(tr) == 0:
            continue
        raw = tr.copy()
        t = tr[:, :6].astype(np.int64)
        t0 = t[:, 0].min()
        if which == 0:
            end = t[:, 5]
            ph = np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4]], 1)
            names = ["gather", "wait", "extend-add", "eliminate", "publish"]
            mm = (tr[:, 7] & np.uint64(0xffffffff)).astype(np.int64)
            accA = (tr[:, 7] >> np.uint64(32)).astype(np.int64)
            accB = (tr[:, 6] >> np.uint64(32)).astype(np.int64)
            tr = tr.copy()
            tr[:, 6] = tr[:, 6] & np.uint64(0xffffffff)
            big = np.argsort(-ph[:, 3])[:4]
            log("      slowest eliminations: " + "; ".join(
                f"m={mm[i]} elim {ph[i, 3] / 1e3:.1f}us (diag {accA[i] / 1e3:.1f}, trsm {accB[i] / 1e3:.1f}, trailing {(ph[i, 3] - accA[i] - accB[i]) / 1e3:.1f})"
                for i in big))
        else:
            end = t[:, 3]
            ph = np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2]], 1)
            names = ["prefetch", "wait", "solve"]
            mm = tr[:, 7].astype(np.int64)
        span = (end.max() - t0) / 1e3
        log(f"    [{label}] {name}: {len(tr)} tasks, span {span:.1f} us; per-task phase medians (us): " +
            ", ".join(f"{n} {np.median(ph[:, i]) / 1e3:.2f}" for i, n in enumerate(names)) +
            "; phase sums (ms): " + ", ".join(f"{n} {ph[:, i].sum() / 1e6:.2f}" for i, n in enumerate(names)))
        dur = (end - t[:, 0])
        top = np.argsort(-dur)[:6]
        log("      longest tasks (us, m): " + ", ".join(f"{dur[i] / 1e3:.1f}/m={mm[i]}" for i in top))
        busy = ph.sum() - ph[:, 1].sum()
        log(f"      non-wait busy time {busy / 1e6:.2f} ms over {len(tr)} tasks -> {busy / 1e3 / len(tr):.2f} us per task")
        if DUMP is not None:
            DUMP[f"{label}:{name}"] = raw


def borrowed_plan(L, param_ptr):
    p = HostPlan.__new__(HostPlan)
    p.L = L
    p.p = C.c_void_p(L.asam_dbg_plan_of_param(param_ptr))
    p.close = lambda: None
    return p


def check_batch(L, d, n, have_ref):
    sub = d.head(n)
    log(f"--- batch n={n} edges={sub.n_edges}")
    h = H.Harness("b200")
    h.load_full(sub)
    c0 = h.chi2()
    ms = h.batch()
    log(f"    chi2_0={c0:.9g} batch wall={ms:.3f} ms, chi2_1={h.chi2():.9g}")
    dev = L.asam_dbg_dev_of_graph(h.graph_ptr())
    plan = borrowed_plan(L, h.param_ptr())
    info = plan.info()
    log("    plan", info)
    # --- kernel 1: Hessian in HBM vs emulation
    S = info["n_slots"]
    Ad = np.zeros((n, 3, 3)); Ao = np.zeros((max(S, 1), 3, 3)); B = np.zeros((n, 3))
    L.asam_debug_read_hessian(dev, n, S, Ad.ctypes.data_as(_dp), Ao.ctypes.data_as(_dp), B.ctypes.data_as(_dp))
    ftype = np.r_[2, np.ones(sub.n_edges, dtype=np.int32)].astype(np.int32)
    fa = np.r_[0, sub.ea].astype(np.int32); fb = np.r_[-1, sub.eb].astype(np.int32)
    fz = np.vstack([[0, 0, 0], sub.ez]); fW = np.vstack([[1e4, 0, 0, 0, 1e4, 0, 0, 0, 1e3], sub.eW])
    node2q = plan.array("node2q"); fslot = plan.array("fslot"); q2node = plan.array("q2node")
    Hm = emul.Hessian(n, S); Hm.reset(n, 1e-4)
    lp = sub.init.copy()
    Hm.linearize(range(len(ftype)), ftype, fa, fb, fz, fW, lp, lp, node2q, fslot)
    triu = np.triu(np.ones((3, 3), bool))
    e_d = np.abs(Ad - Hm.Adiag)[:, triu].max() / max(1.0, np.abs(Hm.Adiag).max())
    e_o = (np.abs(Ao[:S] - Hm.Aoff).max() / max(1.0, np.abs(Hm.Aoff).max())) if S else 0.0
    e_b = np.abs(B - Hm.B).max() / max(1.0, np.abs(Hm.B).max())
    log(f"    k_linearize: rel err Adiag {e_d:.3e} Aoff {e_o:.3e} B {e_b:.3e}")
    # --- kernel 2/3: y, x vs emulation
    fr = emul.Fronts(); fr.ensure(n)
    desc = plan.descs(); ipool = plan.array("ipool")
    emul.factor(fr, Hm, desc, ipool, q2node, plan.array("tasks"), plan.array("nwait"))
    emul.backsolve(fr, desc, ipool, plan.array("btasks"))
    y = np.zeros(3 * n); x = np.zeros(3 * n)
    L.asam_download_y(dev, 0, n, y.ctypes.data_as(_dp)); L.asam_download_x(dev, 0, n, x.ctypes.data_as(_dp))
    e_y = np.abs(y - fr.y[:3 * n]).max() / max(1.0, np.abs(fr.y).max())
    e_x = np.abs(x - fr.x[:3 * n]).max() / max(1.0, np.abs(fr.x).max())
    log(f"    k_factor: rel err y {e_y:.3e}   k_backsolve: rel err x {e_x:.3e}")
    if not (e_y < 1e-8) or not (e_x < 1e-8):
        bad = np.argsort(-np.abs(y - fr.y[:3 * n]))[:5]
        log("    worst y idx", bad, y[bad], fr.y[bad])
        # first bad supernode in task order
        for s in plan.array("tasks"):
            first, cb = int(desc["first"][s]), int(desc["cb"][s])
            seg = slice(3 * first, 3 * (first + cb))
            if np.abs(y[seg] - fr.y[seg]).max() > 1e-8 * max(1, np.abs(fr.y).max()):
                m = 3 * int(desc["mb"][s])
                ld = (m + 2) & ~1  # ASAM_LD
                Fd = np.zeros(ld * m)
                L.asam_debug_read_front(dev, int(desc["f_off"][s]), ld * m, Fd.ctypes.data_as(_dp))
                Fd_m = Fd.reshape(m, ld).T[:m, :]  # column-major, ld = ASAM_LD(m) -> [row, col]
                Fe = fr.F[int(desc["f_off"][s])]
                log(f"    first bad supernode {s}: first={first} cb={cb} mb={desc['mb'][s]} level={desc['level'][s]} "
                    f"ch={desc['ch_cnt'][s]} a_cnt={desc['a_cnt'][s]}")
                log("    front diff (lower) max", np.abs(np.tril(Fd_m) - np.tril(Fe)).max())
                log("    dev front\n", np.array2string(np.tril(Fd_m)[:9, :9], precision=4))
                log("    emu front\n", np.array2string(np.tril(Fe)[:9, :9], precision=4))
                break
    st = h.states()
    if have_ref:
        r = H.Harness("reference"); r.load_full(sub); r.batch()
        d = st - r.states()
        log(f"    states vs reference: max abs diff {np.abs(d).max():.3e}; chi2 {h.chi2():.9g} vs {r.chi2():.9g}")
        r.close()
    h.close()


def check_replay(d, nsteps, have_ref):
    log(f"--- replay lockstep {nsteps} steps")
    if not have_ref:
        log("    (no reference on this box)")
        return
    a = H.Harness("b200"); b = H.Harness("reference")
    a.replay_begin(d); b.replay_begin(d)
    worst = 0.0
    t_a = t_b = 0.0
    for k in range(1, nsteps + 1):
        ca, ma, ia = a.replay_to(k)
        cb, mb, ib = b.replay_to(k)
        t_a += ma[0]; t_b += mb[0]
        sa, sb = a.states(), b.states()
        err = np.abs(sa - sb).max()
        worst = max(worst, err)
        bad = (ia[0][0] != ib[0][0]) or (ia[0][1] != ib[0][1]) or err > 1e-6
        if bad or k in (1, 2, 3, 10, 50, 100) or k == nsteps:
            log(f"    step {k}: naff {ia[0][0]}/{ib[0][0]} start_over {ia[0][1]}/{ib[0][1]} err {err:.3e} "
                f"chi2 {ca[0]:.9g}/{cb[0]:.9g}  ms {ma[0]:.3f}/{mb[0]:.3f}")
        if bad:
            i = int(np.argmax(np.abs(sa - sb).max(axis=1)))
            log("    FIRST MISMATCH at node", i, sa[i], sb[i], "order tail", a.ordering()[-5:], b.ordering()[-5:])
            break
    log(f"    worst abs state diff {worst:.3e}; total ms b200 {t_a:.1f} reference {t_b:.1f}")
    a.close(); b.close()


def timing(L, d):
    log("--- timing M3500 batch (warm plan)")
    h = H.Harness("b200")
    h.load_full(d)
    init = d.init.copy()
    h.batch()
    dev = L.asam_dbg_dev_of_graph(h.graph_ptr())
    L.asam_set_timing(dev, 1)
    walls, ks = [], []
    for _ in range(20):
        h.set_states(init)
        walls.append(h.batch())
        lin = C.c_float(); fac = C.c_float(); bs = C.c_float()
        L.asam_last_kernel_ms(dev, C.byref(lin), C.byref(fac), C.byref(bs))
        ks.append((lin.value, fac.value, bs.value))
    ks = np.array(ks)
    log(f"    wall ms median {np.median(walls):.3f} min {np.min(walls):.3f}; kernels ms median lin {np.median(ks[:,0]):.4f} "
        f"factor {np.median(ks[:,1]):.4f} backsolve {np.median(ks[:,2]):.4f}")
    L.asam_set_trace(dev, 1)
    h.set_states(init)
    h.batch()
    trace_report(L, dev, 8192, "M3500 batch")
    plan = borrowed_plan(L, h.param_ptr())
    DUMP["M3500 batch:desc"] = plan.array("desc")
    L.asam_set_trace(dev, 0)
    h.close()


def replay_profile(L, d, nsteps):
    """Host-phase breakdown of the incremental path over a replay (no reference)."""
    L.asam_dbg_profile.argtypes = [_dp, C.c_int]
    prof = np.zeros(24)
    h = H.Harness("b200")
    h.replay_begin(d)
    h.replay_to(2, want_chi2=False)
    L.asam_dbg_profile(prof.ctypes.data_as(_dp), 1)
    t0 = time.perf_counter()
    _, ms, info = h.replay_to(nsteps, want_chi2=False)
    wall = (time.perf_counter() - t0) * 1e3
    L.asam_dbg_profile(prof.ctypes.data_as(_dp), 1)
    pp = np.zeros(8)
    L.asam_dbg_plan_profile_get.argtypes = [_dp, C.c_int]
    L.asam_dbg_plan_profile_get(pp.ctypes.data_as(_dp), 1)
    log(f"    plan_append internals (ms, whole run incl. warm-up steps): host symbolic {pp[0]:.1f}, asam_reserve {pp[1]:.1f}, uploads {pp[2]:.1f}")
    ninc, nbatch = prof[8], prof[10]
    log(f"--- replay profile: {len(ms)} steps in {wall:.1f} ms (sum of api calls {ms.sum():.1f} ms), inc calls {ninc:.0f}, "
        f"full-traversal steps {prof[17]:.0f}, batch escalations {nbatch:.0f} taking {prof[7]:.1f} ms")
    names = ["pre(sync factors, tree grow, mark)", "plan_append(+uploads)", "linearize+factor launch", "tree append",
             "backsolve launch + x download (GPU wait)", "status download", "apply_solution"]
    for i, nm in enumerate(names):
        log(f"    inc phase {nm}: total {prof[i]:.1f} ms, per call {1e3 * prof[i] / max(ninc, 1):.1f} us")
    bn = ["host gather", "plan (cache check/build)", "uploads+launches", "x download (GPU wait)", "status", "tree+update"]
    for i, nm in enumerate(bn):
        log(f"    batch phase {nm}: total {prof[11 + i]:.1f} ms, per call {prof[11 + i] / max(prof[9], 1):.3f} ms")
    naff = info[:, 0]
    for lo, hi in ((0, 5), (6, 20), (21, 100), (101, 10000)):
        sel = (naff >= lo) & (naff <= hi)
        if sel.any():
            log(f"    steps with naffected in [{lo},{hi}]: {sel.sum()} steps, median {np.median(ms[sel]):.3f} ms, mean {ms[sel].mean():.3f} ms")
    h.close()


def trace_steps(L, d, steps):
    """Kernel traces of individual incremental steps (indices into the demo replay)."""
    h = H.Harness("b200")
    h.replay_begin(d)
    for k in sorted(steps):
        h.replay_to(k, want_chi2=False)          # steps 0..k-1 done
        dev = L.asam_dbg_dev_of_graph(h.graph_ptr())
        L.asam_set_trace(dev, 1)
        L.asam_set_timing(dev, 1)
        _, ms, info = h.replay_to(k + 1, want_chi2=False)
        lin = C.c_float(); fac = C.c_float(); bs = C.c_float()
        L.asam_last_kernel_ms(dev, C.byref(lin), C.byref(fac), C.byref(bs))
        log(f"--- step {k}: naffected {info[0][0]} api {ms[0]:.3f} ms; kernels lin {lin.value:.4f} factor {fac.value:.4f} backsolve {bs.value:.4f} ms")
        plan = borrowed_plan(L, h.param_ptr())
        trace_report(L, dev, 8192, f"step {k}")
        DUMP[f"step {k}:desc"] = plan.array("desc")
        L.asam_set_trace(dev, 0)
    h.close()


def check_synth(L, N, have_ref, iters=2):
    """Synthetic Manhattan graph: exercises the big-front (global-memory) path of k_factor."""
    from aprilsam_b200 import datasets
    d = datasets.manhattan_dense(N, seed=1)
    log(f"--- synthetic manhattan_dense N={N} edges={d.n_edges}")
    h = H.Harness("b200")
    h.load_full(d)
    t0 = time.perf_counter()
    ms = h.batch()
    log(f"    cold batch wall {ms:.1f} ms (python {1e3 * (time.perf_counter() - t0):.1f})")
    dev = L.asam_dbg_dev_of_graph(h.graph_ptr())
    plan = borrowed_plan(L, h.param_ptr())
    log("    plan", plan.info())
    L.asam_set_timing(dev, 1)
    st1 = h.states()
    r = None
    if have_ref and N <= 30000:
        r = H.Harness("reference"); r.load_full(d)
        tr = r.batch()
        err = np.abs(st1 - r.states())
        log(f"    iter 1 vs reference ({tr:.1f} ms): max abs state diff {err.max():.3e}, chi2 {h.chi2():.9g} vs {r.chi2():.9g}")
    for it in range(iters):
        ms = h.batch()
        lin = C.c_float(); fac = C.c_float(); bs = C.c_float()
        L.asam_last_kernel_ms(dev, C.byref(lin), C.byref(fac), C.byref(bs))
        msg = f"    warm batch {it}: wall {ms:.2f} ms; kernels lin {lin.value:.3f} factor {fac.value:.3f} backsolve {bs.value:.3f} ms"
        if r is not None:
            r.batch()
            msg += f"; max abs state diff {np.abs(h.states() - r.states()).max():.3e}"
        log(msg)
    L.asam_set_trace(dev, 1)
    h.batch()
    trace_report(L, dev, 200000, f"manhattan {N}")
    DUMP[f"manhattan {N}:desc"] = plan.array("desc")
    L.asam_set_trace(dev, 0)
    if r is not None:
        r.close()
    h.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--synth", default="", help="comma list of synthetic graph sizes")
    ap.add_argument("--sizes", default="1,2,10,100,3500")
    ap.add_argument("--replay", type=int, default=300)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--profile-replay", type=int, default=0)
    ap.add_argument("--trace-steps", default="")
    args = ap.parse_args()
    d = H.PoseGraphData.load(os.path.join(ROOT, "tests", "golden", "m3500.npz"))
    L = dev_api()
    have_ref = H.available("reference")
    t0 = time.time()
    for n in [int(s) for s in args.sizes.split(",") if s]:
        check_batch(L, d, n, have_ref)
    if args.replay > 0:
        check_replay(d, args.replay, have_ref)
    if args.time:
        timing(L, d)
    if args.profile_replay > 0:
        replay_profile(L, d, args.profile_replay)
    if args.trace_steps:
        trace_steps(L, d, [int(x) for x in args.trace_steps.split(",")])
    for n in [int(s) for s in args.synth.split(",") if s]:
        check_synth(L, n, have_ref)
    if DUMP:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "traces.npz"), **{k.replace(" ", "_").replace(":", "__"): v for k, v in DUMP.items()})
    log(f"done in {time.time() - t0:.1f}s")
