# 100 k batch, sharded over N GPUs of one box: bash tools/scale_100k.sh "2 4 8"
for N in ${1:-1 2 4 8}; do
  if [ $N = 1 ]; then
    timeout 300 python bench.py --workload manhattan_batch --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_100k_batch_n$N.json 2> gpurun_out/r2j_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --workload manhattan_batch --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_100k_batch_n$N.json 2> gpurun_out/r2j_n$N.err
  fi
  tail -1 gpurun_out/r2j_n$N.err | cut -c1-200
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j_100k_batch_n*.json")):
    try:
        lines=[l for l in open(f) if l.startswith("{")]
        j=json.loads(lines[-1]); print(f.split("_n")[-1], "value", round(j["value"],2), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],2), j.get("kernel_ms"))
    except Exception as e: print(f, "ERR", e)
PY
