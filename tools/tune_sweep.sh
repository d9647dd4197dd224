#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
python tools/tune_batch.py --tag default --save /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=0 python tools/tune_batch.py --tag tile0_dfma --check /tmp/base100k.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --tag m3500 >> $out 2>&1
grep TUNE $out
