#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_TILE_MODE=0 python tools/tune_batch.py --tag tile0_dfma --save /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=2 python tools/tune_batch.py --tag tile2_dmma_bulkcols --check /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=3 timeout 120 python tools/tune_batch.py --tag tile3_rowmajor_ws --check /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=0 python tools/tune_batch.py --poses 30000 --tag 30k_tile0 --save /tmp/base30k.npy >> $out 2>&1
ASAM_TILE_MODE=3 timeout 120 python tools/tune_batch.py --poses 30000 --tag 30k_tile3 --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
