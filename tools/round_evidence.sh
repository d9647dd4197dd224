timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
R=r2l
timeout 200 python bench.py > gpurun_out/${R}_m3500_batch.json 2> gpurun_out/${R}.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/${R}_m3500_batch_reference.json 2>> gpurun_out/${R}.err
timeout 300 python bench.py --workload m3500_replay --steps 3490 > gpurun_out/${R}_m3500_replay.json 2>> gpurun_out/${R}.err
timeout 400 python bench.py --workload manhattan_batch --steps 10 --warmup 3 > gpurun_out/${R}_100k_batch.json 2>> gpurun_out/${R}.err
timeout 600 python bench.py --workload manhattan_replay --poses 100000 --replay-from 50000 --steps 300 --warmup 3 > gpurun_out/${R}_100k_replay_from50k.json 2>> gpurun_out/${R}.err
tail -4 gpurun_out/${R}.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.2f"%j["value"], "ms %.4f"%j["ms_per_step"], "e2e %.2f"%j["e2e"]["value"], j.get("kernel_ms"), "cpu", (j.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
bash tools/profile_round.sh r2l launches 2>&1 | tail -12
