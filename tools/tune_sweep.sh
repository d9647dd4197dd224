#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
python tools/tune_batch.py --tag default --save /tmp/base100k.npy >> $out 2>&1
ASAM_STAGED=0 timeout 120 python tools/tune_batch.py --tag nostage --check /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=0 timeout 120 python tools/tune_batch.py --tag tile0 --check /tmp/base100k.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_default >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_default >> $out 2>&1
grep TUNE $out
