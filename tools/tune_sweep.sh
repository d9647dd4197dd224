#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
python tools/tune_batch.py --tag default --save /tmp/base100k.npy >> $out 2>&1
ASAM_PB_SMEM=24 timeout 120 python tools/tune_batch.py --tag pbsmem24 --check /tmp/base100k.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_default --save /tmp/basem.npy >> $out 2>&1
ASAM_PB_SMEM=24 python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_pbsmem24 --check /tmp/basem.npy >> $out 2>&1
ASAM_PB_SMEM=6 python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_pbsmem6 --check /tmp/basem.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_default >> $out 2>&1
grep TUNE $out
