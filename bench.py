#!/usr/bin/env python
"""bench.py -- Gauss-Newton solves/sec of the april_graph_cholesky{,_inc} path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload m3500_batch|manhattan_batch|m3500_replay|manhattan_replay] [--poses P]

One "step" is one pass of the hot path over one batch of synthetic/fixture input:
  *_batch   one april_graph_cholesky() call on the full graph (relinearise -> assemble ->
            factor -> 2 triangular solves -> state update), node states reset to the VERTEX2
            initial estimate before every call so each call does identical work;
  *_replay  one april_graph_cholesky_inc() call of the demo-protocol pose-by-pose replay
            (includes any batch escalation it triggers).
Default workload: BASELINE.json configs[1] = "M3500 batch Cholesky on 1xB200".

JSON line (rank 0):
  value        solves/s with every input already resident in HBM: the GPU pipeline
               (reset + k_linearize + k_factor + k_backsolve) timed per step with CUDA events on
               the library's stream, L2 flushed between steps (outside the timed events)
  e2e.value    solves/s through the public C API (april_graph_cholesky on HOST structs): host
               pose gather, H2D, kernels, D2H of the solution, host state update all inside
  roofline     dominant kernel k_factor: algorithmic bytes 8*nnz(A_upper)+16*nnz(L) per launch
                / measured launch time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the reference's own CPU implementation (oracle/_ref) on this box, 1 thread
Multi-GPU (one process per GPU, torchrun): M3500 and every incremental workload do not shard
("replicas only"): each rank solves its own replica, value = total solves of
all ranks / max rank time, scaling "weak".  The 100 k batch workload (manhattan_batch) shards ONE
solve over the GPUs: elimination-tree shards per rank, NCCL broadcast of the shard roots' update
matrices and of the solution segments, the top of the tree replicated; every
rank calls april_graph_cholesky on its copy of the graph, value = solves / max rank time, scaling
"strong".  --shard on|off overrides.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from aprilsam_b200 import capi, datasets  # noqa: E402
from aprilsam_b200 import harness as H  # noqa: E402

METRIC = "Gauss-Newton solves/sec (batch + incremental) on M3500 & 100k-pose graph"


def load_workload(name: str, poses: int):
    if name.startswith("m3500"):
        d = H.PoseGraphData.load(os.path.join(ROOT, "tests", "golden", "m3500.npz"))
        label = "M3500 (3500 poses, 5453 xyt + 1 xytpos factors)"
    elif name == "manhattan_batch":
        d = datasets.manhattan_dense(poses, seed=1)
        label = f"synthetic Manhattan seed 1, {d.n_nodes} poses, {d.n_edges} xyt + 1 xytpos factors"
    elif name == "manhattan_replay":
        d = datasets.manhattan_sparse(poses, seed=1)
        label = f"synthetic Manhattan 5% closures seed 1, {d.n_nodes} poses, {d.n_edges} xyt + 1 xytpos factors"
    else:
        raise SystemExit(f"unknown workload {name}")
    return d, label


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    def __init__(self, device: int, period_ms: int = 500):
        self.p = None
        self.device = device
        self.period_ms = period_ms
        self.path = os.path.join("/tmp", f"asam_clocks_{os.getpid()}.csv")

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                                       "-i", str(self.device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 9:
                    continue
                sm.append(float(t[1]))
                mx.append(float(t[2]))
                for nm, v in zip(names, t[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def dist_setup(ngpus: int, backend: str | None = None):
    """One process per GPU (torchrun env).  backend "gloo" is used by the CPU tests."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    backend = backend or os.environ.get("ASAM_BENCH_BACKEND", "nccl")
    if world > 1:
        import torch
        import torch.distributed as dist
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
        return world, rank, local, dist
    return 1, 0, local, None


def barrier_max(dist, local, value: float) -> float:
    """Barrier, then the maximum of `value` over all ranks (device-side for NCCL)."""
    if dist is None:
        return value
    import torch
    cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", local) if cuda else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if cuda:
        torch.cuda.synchronize()
    return float(t.item())


def aggregate_rate(dist, local, world: int, steps_this_rank: int, seconds_this_rank: float) -> float:
    """Whole-job rate: every rank runs `steps_this_rank` steps of its own replica; the job takes as
    long as the slowest rank (replicas only)."""
    return world * steps_this_rank / barrier_max(dist, local, seconds_this_rank)


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ----------------------------------------------------------------------------------------------
def time_reference_batch(d, calls: int, warm: int = 1):
    h = H.Harness("reference")
    h.load_full(d)
    init = d.init.copy()
    ms = []
    for i in range(warm + calls):
        h.set_states(init)
        t = h.batch()
        if i >= warm:
            ms.append(t)
    h.close()
    return np.array(ms)


BULK_START = 2000  # replay windows that start later are reached by one batch solve, not by replay


def replay_seek(h, d, s_begin: int):
    """Bring harness `h` to the state "first s_begin poses solved".  Small offsets replay the demo
    protocol from pose 0; large ones (a full CPU replay of 100 k poses takes
    hours) load the first s_begin poses at once, start them at the generator's ground truth (what a
    replay from pose 0 would have tracked: dead-reckoning over 50 k poses is metres off and
    Gauss-Newton diverges from there, the linear systems become numerically meaningless) and run
    two batch solves -- identical for both arms."""
    h.replay_begin(d)
    if s_begin <= BULK_START:
        h.replay_to(s_begin, want_chi2=False)
    else:
        sub = d.head(s_begin)
        h.load_full(sub)
        if sub.truth is not None:
            h.set_states(sub.truth)
        h.batch()
        h.batch()


def time_reference_replay(d, s_begin: int, steps: int):
    h = H.Harness("reference")
    replay_seek(h, d, s_begin)
    _, ms, _ = h.replay_to(s_begin + steps, want_chi2=False)
    h.close()
    return ms


def run_reference(args, d, label, world, rank):
    if rank != 0:
        return
    if not H.available("reference"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    if args.workload.endswith("_batch"):
        ms = time_reference_batch(d, args.steps, args.warmup)
        sample = f"{args.steps} april_graph_cholesky calls on the full graph after {args.warmup} warm-up"
    else:
        s0 = replay_start(args, d)
        ms = time_reference_replay(d, s0, args.steps)
        sample = f"demo replay steps [{s0}, {s0 + len(ms)})"
    total_s = float(ms.sum()) / 1e3
    val = len(ms) / total_s
    line = {"metric": METRIC, "value": val, "unit": "solves/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": int(len(ms)), "warmup": args.warmup, "ms_per_step": float(ms.mean()), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic" if "manhattan" in args.workload else "fixture M3500 (public dataset) + synthetic prior",
            "config": {"workload": args.workload, "graph": label},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def replay_start(args, d) -> int:
    return max(1, min(args.replay_from, d.n_nodes - 1))


# ----------------------------------------------------------------------------------------------
# b200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args, d, label, world, rank, local, dist):
    L = capi.lib()
    if L.asam_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    hbm_peak, peak_src = peaks()
    is_batch = args.workload.endswith("_batch")
    sampler = ClockSampler(local, period_ms=int(os.environ.get("ASAM_CLOCK_PERIOD_MS", "500")))

    sharded = world > 1 and is_batch and (args.shard == "on" or (args.shard == "auto" and args.workload == "manhattan_batch"))
    if sharded:
        capi.comm_init_torch(dist, local)
        capi.check(L.asam_comm_set_sharding(1), "asam_comm_set_sharding")
    h = H.Harness("b200")
    init = d.init.copy()
    if is_batch:
        h.load_full(d)
        t0 = time.perf_counter()
        h.batch()  # cold call: ordering + symbolic analysis + plan upload
        cold_ms = (time.perf_counter() - t0) * 1e3
    else:
        s0 = replay_start(args, d)
        replay_seek(h, d, s0)
        cold_ms = None
    dev = L.asam_dbg_dev_of_graph(h.graph_ptr())
    pinfo = capi.plan_info(L.asam_dbg_plan_of_param(h.param_ptr()))
    L.asam_set_timing(dev, 1)

    # ---- e2e: through the public API, host structs in, host structs out ----------------------
    K, W = args.steps, args.warmup
    e2e_ms, kern = [], []
    launches0 = h2d0 = d2h0 = 0
    if rank == 0:
        sampler.start()
    if is_batch:
        for i in range(W + K):
            h.set_states(init)
            if i == W:
                barrier_max(dist, local, 0.0)
                launches0, h2d0, d2h0 = capi.counters(dev)
            e2e_t = h.batch()
            if i >= W:
                e2e_ms.append(e2e_t)
                kern.append(capi.kernel_ms(dev))
    else:
        h.replay_to(s0 + W, want_chi2=False)
        barrier_max(dist, local, 0.0)
        launches0, h2d0, d2h0 = capi.counters(dev)
        _, ms, info = h.replay_to(s0 + W + K, want_chi2=False)
        e2e_ms = list(ms)
    launches1, h2d1, d2h1 = capi.counters(dev)
    nsteps = len(e2e_ms)
    # replicas: every rank solved its own copy; sharded: all ranks worked on the SAME solves
    jobs = 1 if sharded else world
    e2e_val = aggregate_rate(dist, local, jobs, nsteps, float(np.sum(e2e_ms)) / 1e3)

    # ---- device-resident: same pipeline, inputs already in HBM, CUDA events per step ----------
    dev_ms = []
    fac_ms = []
    if is_batch:
        h.set_states(init)
        h.batch()
        N, S, F = pinfo["N"], pinfo["n_slots"], pinfo["n_factors"]
        ms = C.c_float()
        launches_a = capi.counters(dev)[0]
        for i in range(W + K):
            if i == W:
                launches_a = capi.counters(dev)[0]
            capi.check(L.asam_l2_flush(dev), "l2_flush")
            capi.check(L.asam_timer_start(dev), "timer")
            capi.check(L.asam_hessian_reset(dev, N, S, N, 1e-4), "reset")
            capi.check(L.asam_linearize(dev, 0, F, None), "linearize")
            capi.check(L.asam_factor_full(dev), "factor")
            capi.check(L.asam_backsolve_full(dev), "backsolve")
            capi.check(L.asam_timer_stop(dev, C.byref(ms)), "timer")
            if i >= W:
                dev_ms.append(ms.value)
                fac_ms.append(capi.kernel_ms(dev)[1])
        launches_dev = capi.counters(dev)[0] - launches_a
        st = C.c_int()
        L.asam_factor_status(dev, C.byref(st))
        if st.value != 0:
            raise SystemExit(f"bench.py: factorisation status {st.value}")
        value = aggregate_rate(dist, local, jobs, len(dev_ms), float(np.sum(dev_ms)) / 1e3)
        ms_per_step = float(np.mean(dev_ms))
        gpu_launches = launches_dev
    else:
        value = e2e_val
        ms_per_step = float(np.mean(e2e_ms))
        gpu_launches = launches1 - launches0
    clocks = sampler.stop() if rank == 0 else {}

    # ---- roofline of the dominant kernel (k_factor) -------------------------------------------
    nnz_l = 9 * (pinfo["nnz_l_blocks"] - pinfo["N"]) + 6 * pinfo["N"]
    nnz_a = 6 * pinfo["N"] + 9 * pinfo["n_slots"]
    alg_bytes = 8 * nnz_a + 16 * nnz_l
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(args.workload, {}).get("bytes")
    except Exception:
        pass
    if is_batch and fac_ms:
        t_fac = float(np.mean(fac_ms)) * 1e-3
        achieved = alg_bytes / t_fac / 1e9
        fp64_peak = 37.1e3  # GFLOP/s, measured on this pool (profiles/r2_fp64_pipe_ubench.txt); tcgen05 has no FP64 kind
        roof = {"kernel": "k_factor (+k_factor_leaf on large graphs)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": t_fac * 1e3,
                "fp64_gflops": pinfo["flops"] / t_fac / 1e9, "flops_per_launch": pinfo["flops"],
                "fp64_peak_gflops": fp64_peak, "fp64_frac": pinfo["flops"] / t_fac / 1e9 / fp64_peak,
                "note": "latency-bound: the dependent chain of supernode levels / panel steps, not bytes or flops, sets the time"}
    else:
        roof = {"kernel": "k_factor", "bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                "traffic": None, "peak_source": peak_src}

    # ---- cpu baseline (rank 0, N=1): the reference on this box's host cores -------------------
    cpu = None
    if rank == 0 and world == 1 and H.available("reference") and not args.no_cpu_baseline:
        if is_batch:
            probe = time_reference_batch(d, 1, 0)
            calls = int(max(1, min(300, 10e3 / max(probe[0], 1e-3))))
            ms = time_reference_batch(d, calls, 0)
            cpu = {"value": calls / (float(ms.sum()) / 1e3), "unit": "solves/s", "cores": 1, "kind": "reference",
                   "sample": f"{calls} april_graph_cholesky calls of the same graph (oracle/_ref, 1 thread: the reference has no threads)"}
        else:
            ms = time_reference_replay(d, s0, W + (K if d.n_nodes <= 5000 else min(K, 1000)))[W:]
            cpu = {"value": len(ms) / (float(ms.sum()) / 1e3), "unit": "solves/s", "cores": 1, "kind": "reference",
                   "sample": f"replay steps [{s0 + W}, {s0 + W + len(ms)}) (oracle/_ref, deterministic clock)"
                             + ("" if s0 <= BULK_START else f"; first {s0} poses loaded at once at ground truth + two batch solves")}
    h.close()

    if rank != 0:
        return
    per_step = max(nsteps, 1)
    line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic" if "manhattan" in args.workload else "fixture M3500 (public dataset) + synthetic prior",
            "config": {"workload": args.workload, "graph": label,
                       "parallelism": (f"elimination-tree shards x{world}, top of the tree replicated, NCCL broadcast of shard "
                                       f"root fronts and solution segments") if sharded else f"replicas x{world}",
                       "cache": "L2 flushed (384 MiB overwrite) between timed steps" if is_batch else "working set grows each step",
                       "plan": {k: pinfo[k] for k in ("N", "nsn", "n_slots", "n_levels", "max_m", "nnz_l_blocks")},
                       "cold_first_call_ms": cold_ms},
            "e2e": {"value": e2e_val, "unit": "solves/s", "ms_per_step": float(np.mean(e2e_ms)),
                    "h2d_bytes_per_step": (h2d1 - h2d0) // per_step, "d2h_bytes_per_step": (d2h1 - d2h0) // per_step},
            "gpu_launches": int(gpu_launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks}
    if kern:
        k = np.array(kern)
        line["kernel_ms"] = {"k_linearize": float(np.median(k[:, 0])), "k_factor": float(np.median(k[:, 1])),
                             "k_backsolve": float(np.median(k[:, 2]))}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="m3500_batch",
                    choices=["m3500_batch", "manhattan_batch", "m3500_replay", "manhattan_replay"])
    ap.add_argument("--poses", type=int, default=100000)
    ap.add_argument("--replay-from", type=int, default=1, help="replay workloads: first timed step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", choices=["auto", "on", "off"], default="auto",
                    help="--gpus N > 1: shard the elimination tree of ONE solve over the GPUs (NCCL) instead of "
                         "running N replicas; auto = on for manhattan_batch (SURVEY.md section 8e)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200" and args.workload.endswith("_batch"):
        args.warmup = 3
    d, label = load_workload(args.workload, args.poses)
    if args.impl == "reference":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, d, label, world, rank)
        return
    world, rank, local, dist = dist_setup(args.gpus)
    try:
        run_b200(args, d, label, world, rank, local, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
