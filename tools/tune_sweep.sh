#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_STAGED=0 ASAM_MID=0 python tools/tune_batch.py --tag nostage_nomid --save /tmp/base100k.npy >> $out 2>&1
ASAM_STAGED=1 ASAM_MID=0 timeout 120 python tools/tune_batch.py --tag staged_nomid --check /tmp/base100k.npy >> $out 2>&1
ASAM_STAGED=0 ASAM_MID=1 timeout 120 python tools/tune_batch.py --tag nostage_mid --check /tmp/base100k.npy >> $out 2>&1
timeout 120 python tools/tune_batch.py --tag default_staged_mid --check /tmp/base100k.npy >> $out 2>&1
ASAM_TASK_ORDER=cp timeout 120 python tools/tune_batch.py --tag default_cp --check /tmp/base100k.npy >> $out 2>&1
ASAM_STAGED=0 ASAM_MID=0 python tools/tune_batch.py --poses 30000 --tag 30k_nostage_nomid --save /tmp/base30k.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_default --check /tmp/base30k.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_default >> $out 2>&1
ASAM_TASK_ORDER=level python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_level >> $out 2>&1
grep TUNE $out
