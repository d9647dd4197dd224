#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_TASK_ORDER=level ASAM_BS_SPLIT=0 python tools/tune_batch.py --tag level_nosplit --save /tmp/base100k.npy >> $out 2>&1
ASAM_BS_SPLIT=0 timeout 120 python tools/tune_batch.py --tag cp_nosplit --check /tmp/base100k.npy >> $out 2>&1
ASAM_TASK_ORDER=level timeout 120 python tools/tune_batch.py --tag level_split --check /tmp/base100k.npy >> $out 2>&1
timeout 120 python tools/tune_batch.py --tag default_cp_split --check /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=0 timeout 120 python tools/tune_batch.py --tag default_tile0 --check /tmp/base100k.npy >> $out 2>&1
ASAM_TASK_ORDER=level python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_level --save /tmp/basem.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_cp --check /tmp/basem.npy >> $out 2>&1
ASAM_TASK_ORDER=level python tools/tune_batch.py --poses 30000 --tag 30k_level --save /tmp/base30k.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_cp --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
