#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_TASK_ORDER=level python tools/tune_batch.py --tag level --save /tmp/base100k.npy >> $out 2>&1
ASAM_TASK_ORDER=cp timeout 120 python tools/tune_batch.py --tag cp --check /tmp/base100k.npy >> $out 2>&1
timeout 120 python tools/tune_batch.py --tag sim --check /tmp/base100k.npy >> $out 2>&1
ASAM_TEAM_MIN=2 timeout 120 python tools/tune_batch.py --tag sim_min2 --check /tmp/base100k.npy >> $out 2>&1
ASAM_TEAM_ROOM=148 timeout 120 python tools/tune_batch.py --tag sim_room148 --check /tmp/base100k.npy >> $out 2>&1
ASAM_TASK_ORDER=level python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_level --save /tmp/basem.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_sim --check /tmp/basem.npy >> $out 2>&1
ASAM_TASK_ORDER=level python tools/tune_batch.py --poses 30000 --tag 30k_level --save /tmp/base30k.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_sim --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
