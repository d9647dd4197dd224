#!/usr/bin/env python
"""profiles/README.md from the committed evidence of a round (no GPU needed).

    python tools/make_profiles_readme.py rd2 > profiles/README.md
"""
import csv
import json
import os
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "rd2"
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def load(name):
    p = os.path.join(P, name)
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def ncu(name):
    p = os.path.join(P, name)
    if not os.path.exists(p):
        return []
    rows = list(csv.reader(open(p)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append({h: (r[i], units[i]) for i, h in enumerate(hdr)})
    return out


def f(x, n=1):
    return "—" if x is None else f"{x:.{n}f}"


b1, ref = load(f"{R}_bench_n1.json"), load(f"{R}_bench_n1_reference.json")
bn = {n: load(f"{R}_bench_n{n}.json") for n in (2, 4, 8)}
print(f"""# profiles/ — measured evidence

Everything here was taken on B200s (148 SMs, 1965 MHz max SM clock) through `gpurun`; nothing was timed under a
profiler except the ncu files themselves.  Files of THIS round (round 2) start with `{R}` (`{R}a_*` = the same set
taken mid-round, before the last kernel changes; `rd2_*_panel_trace.log` = device-stamp traces of the panel step
as it evolved); `r1*` / `r2*` files are round 1's.  `tools/round_evidence.sh` regenerates the set,
`tools/make_profiles_readme.py` this file.

## Files of this round

| file | what |
|---|---|
| `{R}_bench_n1.json`, `{R}_bench_n1_reference.json` | the default `bench.py` line (ALL FOUR workloads, 100 k batch as headline) of both arms on one box, `--steps 20 --warmup 5` |
| `{R}_bench_n1_after_plan_changes.json`, `{R}_plan_profile.log` | the b200 arm again after the last host-side changes (cold plan build), and the phases of the plan build on the GPU box's host with 1 / 8 / 16 threads |
| `{R}_bench_n2.json`, `{R}_bench_n4.json`, `{R}_bench_n8.json` | the same line under `torchrun` on 2 / 4 / 8 GPUs: the 100 k batch solve SHARDED over the GPUs (in-line parity against the single-GPU solve), the other workloads as replicas |
| `{R}_replay_from10k.json`, `{R}_replay_from90k.json` | sparse 100 k replay windows starting at pose 10 000 / 90 000 (the default line has the 50 000 window) |
| `{R}_m3500_batch_launches.csv`, `{R}_100k_batch_launches.csv`, `{R}_m3500_replay_launches.csv` | every kernel launch with its device time (`ncu --metrics gpu__time_duration.sum`, cold cache, serialised: shares, not absolutes) |
| `{R}_m3500_batch_prof_raw.csv`, `{R}_100k_batch_prof_raw.csv` | `--page raw` export of one `ncu --set full` capture of every solve kernel of one step; `ncu_traffic.json` = their DRAM bytes per launch (`tools/ncu_extract.py`), which `bench.py` copies into `roofline_kernels[].traffic` |
| `{R}_sass_excerpt.txt` | `cuobjdump -sass` of the built library: instruction counts per kernel + excerpt (`DMMA.8x8x4`, `UBLKCP.S.G`, `SYNCS.ARRIVE.TRANS64`, `SYNCS.PHASECHK…TRYWAIT`) |
| `{R}_ozaki_study.log` | host-side numerical study (`tools/ozaki_study.py`): node-state error of the batch step when every trailing product is assembled from s int8 slices — 1e-6 needs 6 slices (21 products) on M3500, 7 (28) on an 8000-pose dense world |
| `{R}_memcheck.log` | `compute-sanitizer --tool memcheck` over M3500 batch, 60 replay steps (k_step) and a 12 k-pose synthetic world (leaf kernels, team path with bulk copies): 0 errors |
| `{R}_tune_*.log` | the A/B sweeps behind the defaults (tile modes, staged publish, task order, team sizes, back-solve split, tensor-pipe variants) — same box within a file |
| `rd2_*_panel_trace.log`, `{R}_panel_trace.log`, `rd2_panel_trace_final_ab.log` | per-panel device stamps of the root front: 32.3 µs/panel at the start of the round → 26.2 µs (final build: `{R}_panel_trace.log`) |
| `{R}_step_profile.log` | device stamps of `k_step` (the one-launch small incremental step) over the M3500 replay + host-side phases |
""")
if b1 and ref:
    print("## Headline numbers (one box, both arms back to back; reference = unmodified AprilSAM, 1 thread — it has none)\n")
    print("| workload | reference CPU | aprilsam_b200 e2e (host structs in/out) | device-resident | e2e ÷ reference | e2e uncached (plan rebuilt every call) |")
    print("|---|---|---|---|---|---|")
    rows = [("Manhattan 100 k batch (headline)", b1, ref)]
    for n in ("m3500_batch", "m3500_replay", "manhattan_replay"):
        rows.append((n, b1["workloads"][n], ref["workloads"][n]))
    for name, w, r in rows:
        unc = w.get("e2e_uncached")
        print(f"| {name} | {r['value']:.3f} solves/s | **{w['e2e']['value']:.1f} solves/s** ({w['e2e'].get('ms_per_step', 1e3 / w['e2e']['value']):.3f} ms) | "
              f"{w['value']:.1f} solves/s ({w['ms_per_step']:.3f} ms) | {w['e2e']['value'] / r['value']:.1f}× | "
              + (f"{unc['value']:.2f} solves/s ({unc['ms_per_step']:.1f} ms, plan {unc['plan_build_ms_per_call']:.1f} ms) → {unc['value'] / r['value']:.1f}×" if unc else "—") + " |")
    b1b = load(f"{R}_bench_n1_after_plan_changes.json")
    if b1b and b1b.get("e2e_uncached"):
        u0, u1 = b1["e2e_uncached"], b1b["e2e_uncached"]
        print(f"\nAfter the last host-side changes of the round (cold plan: no per-column sort, symbolic loops on host threads) "
              f"the same line on another box (`{R}_bench_n1_after_plan_changes.json`): plan {u0['plan_build_ms_per_call']:.0f} → "
              f"{u1['plan_build_ms_per_call']:.0f} ms per uncached call, uncached 100 k batch {u0['value']:.2f} → {u1['value']:.2f} solves/s; "
              f"device-resident {b1b['value']:.1f} solves/s, e2e {b1b['e2e']['value']:.1f} solves/s (unchanged within box-to-box spread).")
    print()
    print("Round 1 → round 2, same workloads (device-resident step): 100 k batch 8.67 → "
          f"{b1['ms_per_step']:.2f} ms (k_factor 7.21 → {b1['kernel_ms']['k_factor']:.2f}, k_backsolve 1.37 → {b1['kernel_ms']['k_backsolve']:.2f}); "
          f"M3500 batch 0.637 → {b1['workloads']['m3500_batch']['ms_per_step']:.3f} ms.\n")
    print("### Incremental steps by size (median µs per `april_graph_cholesky_inc` call; same steps, same box)\n")
    print("| workload | bucket | steps | aprilsam_b200 | reference CPU |")
    print("|---|---|---|---|---|")
    for n in ("m3500_replay", "manhattan_replay"):
        w, r = b1["workloads"][n], ref["workloads"][n]
        rb = (w.get("cpu_baseline") or {}).get("latency_by_bucket") or r.get("latency_by_bucket") or {}
        for bk, v in w["latency_by_bucket"].items():
            rv = rb.get(bk, {})
            print(f"| {n} | {bk} | {v['steps']} | {f(v['median_us'])} | {f(rv.get('median_us'))} ({rv.get('steps', 0)} steps) |")
    fs = b1["workloads"]["m3500_replay"].get("fused_small_steps")
    if fs and fs.get("mean_us"):
        print("\n`k_step` (one launch per small step), mean µs per phase on the M3500 replay: "
              + ", ".join(f"{k} {v:.1f}" for k, v in fs["mean_us"].items()) + ".")
    wc = b1["workloads"]["m3500_replay"].get("cpu_baseline_wallclock")
    if wc:
        print(f"\nThe reference AS SHIPPED (wall-clock escalation heuristic active, non-deterministic) runs the M3500 replay at "
              f"{wc['value']:.0f} solves/s on the same box (deterministic clock: {b1['workloads']['m3500_replay']['cpu_baseline']['value']:.0f}); "
              f"aprilsam_b200: {b1['workloads']['m3500_replay']['value']:.0f}.")
    print("\n## Per-kernel roofline (live CUDA-event times inside the e2e calls; peaks: MEASURED_PEAKS.json hbm_gbs, FP64 measured live)\n")
    print("| workload | kernel | algorithmic bytes/launch | time | achieved | of HBM peak | DRAM traffic (ncu) | FP64 |")
    print("|---|---|---|---|---|---|---|---|")
    for name, w in (("100 k batch", b1), ("M3500 batch", b1["workloads"]["m3500_batch"])):
        for e in w["roofline_kernels"]:
            fp = f"{e['fp64_tflops']:.2f} TFLOP/s = {100 * e['fp64_frac']:.1f} % of {e['fp64_peak_tflops']:.1f}" if e.get("fp64_tflops") else "—"
            tr = f"{e['traffic'] / 1e6:.0f} MB" if e.get("traffic") else "—"
            print(f"| {name} | `{e['kernel'].split(' ')[0]}` | {e['algorithmic_bytes_per_launch'] / 1e6:.1f} MB | {e['avg_launch_ms'] * 1e3:.0f} µs | "
                  f"{e['achieved']:.1f} GB/s | {100 * e['frac']:.2f} % | {tr} | {fp} |")
    print()
w10, w90 = load(f"{R}_replay_from10k.json"), load(f"{R}_replay_from90k.json")
if b1 and w10 and w90:
    print("### Sparse 100 k replay by window (solves/s over the timed steps, escalations included)\n")
    print("| first timed pose | aprilsam_b200 |")
    print("|---|---|")
    print(f"| 10 000 | {w10['value']:.0f} |")
    print(f"| 50 000 (default line) | {b1['workloads']['manhattan_replay']['value']:.0f} |")
    print(f"| 90 000 | {w90['value']:.0f} |")
    print()
sc = [(1, b1)] + [(n, bn[n]) for n in (2, 4, 8) if bn[n]]
if len(sc) > 1:
    print("## 100 k batch sharded over the GPUs of one box (strong scaling; device-resident value = solves / max rank time)\n")
    print("| GPUs | device-resident | ms/step | speed-up | e2e | k_factor (incl. exchange) | k_backsolve | parity vs single GPU |")
    print("|---|---|---|---|---|---|---|---|")
    for n, b in sc:
        par = (b.get("parity") or {}).get("sharded_vs_single_gpu_max_rel")
        print(f"| {n} | {b['value']:.1f} solves/s | {b['ms_per_step']:.2f} | {b['value'] / b1['value']:.2f}× | {b['e2e']['value']:.1f} solves/s | "
              f"{b['kernel_ms']['k_factor']:.2f} ms | {b['kernel_ms']['k_backsolve']:.2f} ms | {'—' if par is None else f'{par:.1e}'} |")
    print("\nThe replica workloads of the same lines (value = total solves of all ranks / max rank time):\n")
    print("| GPUs | M3500 batch | M3500 replay | 100 k sparse replay |")
    print("|---|---|---|---|")
    for n, b in sc:
        w = b["workloads"]
        print(f"| {n} | {w['m3500_batch']['e2e']['value']:.0f} | {w['m3500_replay']['value']:.0f} | {w['manhattan_replay']['value']:.0f} |")
    print()
for wl, fn in (("100 k batch", f"{R}_100k_batch_prof_raw.csv"), ("M3500 batch", f"{R}_m3500_batch_prof_raw.csv")):
    rows = ncu(fn)
    if not rows:
        continue
    print(f"## `ncu --set full`, {wl} (one step)\n")
    print("| kernel | duration | grid × block, regs | DRAM read / write | DRAM busy | FP64 pipe active | warps active |")
    print("|---|---|---|---|---|---|---|")
    for r in rows:
        g = lambda k: r.get(k, ("", ""))
        print(f"| `{g('Kernel Name')[0].split('(')[0]}` | {g('gpu__time_duration.sum')[0]} {g('gpu__time_duration.sum')[1]} | "
              f"{g('launch__grid_size')[0]} × {g('launch__block_size')[0]}, {g('launch__registers_per_thread')[0]} | "
              f"{float(g('dram__bytes_read.sum')[0]):.1f} {g('dram__bytes_read.sum')[1]} / {float(g('dram__bytes_write.sum')[0]):.1f} {g('dram__bytes_write.sum')[1]} | "
              f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[0]):.1f} % | "
              f"{float(g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')[0]):.1f} % | "
              f"{float(g('sm__warps_active.avg.pct_of_peak_sustained_active')[0]):.1f} % |")
    print()
print("""## Reading the numbers

* **The factorisation is bound by dependent chains and SM occupancy, not by bytes or flops** (DRAM busy 3 %, FP64 pipe
  active 2 % at 100 k).  Device traces (`tools/panel_trace.py --dump-trace`, analysed in DESIGN.md §4/§6): with the
  simulated ticket order the 148 CTAs are saturated for ≈ 4.2 ms — every front is a short chain of L2 round trips and
  barrier-separated panel steps, one CTA per SM (CTA-time by class at 100 k: shared-memory fronts 27 %, one-CTA team
  fronts 14 %, teams of 2-15 22 %, larger teams 37 %) — followed by a ≈ 0.9 ms tail that is ONE chain: the merged root
  front's 30 panel steps (26.2 µs each: diagonal tile 3.2 + staged 48×48 factorisation 16.4 + the crew's last row-solve
  stage and the team barrier 6.6).  The 48×48 block itself (micro-benchmark `tools/ubench/diag_block.cu`, 11.4 µs
  alone) is 48 chained reciprocal square roots of 131 cycles with two block barriers and three shared-memory round
  trips per 3×3 step: neither a publisher warp nor register-resident 12×12 blocks shortened it (DESIGN.md §4).
* DRAM traffic of `k_factor` is ≈ 3× the algorithmic bytes: a multifrontal method writes every update matrix (Schur
  complement) once and reads it once (1.5 GB of fronts at 100 k against 474 MB of L + A), plus the zero-fill of the
  fronts and the row-major panel workspace of the team path.  It is not the limiter.
* What moved this round (same-box A/B in the `*_tune_*.log` files): extend-add read-modify-writes pipelined (asm phase of
  team fronts 60-75 → 38-42 µs, k_factor 7.2 → 6.6 ms), tensor-pipe tiles with two bulk copies per tile (256×64×48 tile
  14.3 → 4.2 µs; alone: no change of the total, the panel chain hides it), staged publish of the diagonal block
  (6.08 → 5.70 ms), simulated ticket order (30 k: 3.67 → 3.08 ms; M3500: chain-length order 0.457 → 0.426 ms),
  back-substitution one block per CTA for wide supernodes (1.39 → 1.02 ms) and tickets by modelled chain time
  (1.018 → 0.987 ms), team-sized chain fronts merged up to 20 % extra rows (5.85 → 5.71 ms), destination maps of all
  children built while a front waits (M3500 0.441 → 0.432 ms), tuning switches moved from `__device__` globals to
  constant memory (a global load sat on the 3×3 step chain: diagonal block 18.0 → 16.4 µs).  Tried and dropped with numbers: one CTA per
  mid-size front out of HBM (worse from m > 240), a separate two-CTAs-per-SM kernel for fronts ≤ 117 (+0.3 ms: the extra
  launch boundary), several tiles per team worker (6.4 → 6.6-7.0 ms), tensor-pipe update for 12-column shared-memory
  panels (M3500 0.435 → 0.459 ms), nested-dissection ordering (CPU study, DESIGN.md §1).""")
