#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_LEAF_MAX_M=48 python tools/tune_batch.py --workload sparse --tag sparse_leaf48 --save /tmp/bases.npy >> $out 2>&1
python tools/tune_batch.py --workload sparse --tag sparse_default --check /tmp/bases.npy >> $out 2>&1
python tools/tune_batch.py --tag dense100k_default >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_default >> $out 2>&1
grep TUNE $out
