#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_LEAF_TINY=0 python tools/tune_batch.py --tag leaf_one_kernel --save /tmp/base100k.npy >> $out 2>&1
timeout 120 python tools/tune_batch.py --tag leaf_tiny_first --check /tmp/base100k.npy >> $out 2>&1
ASAM_LEAF_TINY=0 python tools/tune_batch.py --poses 30000 --tag 30k_leaf_one --save /tmp/base30k.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_leaf_tiny --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
