#!/bin/bash
# A/B of an environment setting on the two replay workloads + the 30 k batch (run under gpurun)
# usage: replay_ab.sh <log name> VAR=value
out=gpurun_out/${1:-replay_ab}.log
: > $out
for cfg in "" "$2"; do
  for wl in m3500_replay manhattan_replay; do
    echo "== [$cfg] $wl" >> $out
    env $cfg timeout 600 python bench.py --workload $wl --no-cpu-baseline 2>> $out | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    j=json.loads(line)
    keep={k:j.get(k) for k in ('metric','value','unit','ms_per_step','e2e')}
    w=j.get('workloads') or {}
    print(json.dumps(keep)); print(json.dumps({k:{kk:vv for kk,vv in v.items() if kk in ('value','unit','e2e','step_buckets','steps','escalations')} for k,v in w.items()})[:1500])
" >> $out 2>&1
  done
  env $cfg python tools/tune_batch.py --poses 30000 --tag "30k[$cfg]" >> $out 2>&1
done
cat $out
