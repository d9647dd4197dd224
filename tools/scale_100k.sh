for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 300 python bench.py --workload manhattan_batch --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_100k_batch_n$N.json 2> gpurun_out/r2g_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --workload manhattan_batch --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_100k_batch_n$N.json 2> gpurun_out/r2g_n$N.err
  fi
  tail -2 gpurun_out/r2g_n$N.err | cut -c1-300
done
python - <<'PY'
import json
for N in (1,2,4,8):
    try:
        lines=[l for l in open(f"gpurun_out/r2g_100k_batch_n{N}.json") if l.startswith("{")]
        j=json.loads(lines[-1]); print(N, "value", round(j["value"],2), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],2), j.get("kernel_ms"), j["config"]["parallelism"][:40])
    except Exception as e: print(N, "ERR", e)
PY
