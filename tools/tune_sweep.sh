#!/bin/bash
# A/B of the team-path tile implementations and other knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_TILE_MODE=0 python tools/tune_batch.py --tag tile0_dfma --save /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=1 python tools/tune_batch.py --tag tile1_dmma --check /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=2 python tools/tune_batch.py --tag tile2_dmma_bulk --check /tmp/base100k.npy >> $out 2>&1
ASAM_TILE_MODE=0 python tools/tune_batch.py --workload m3500 --tag m3500_tile0 --save /tmp/basem.npy >> $out 2>&1
ASAM_TILE_MODE=2 python tools/tune_batch.py --workload m3500 --tag m3500_tile2 --check /tmp/basem.npy >> $out 2>&1
ASAM_TILE_MODE=0 python tools/tune_batch.py --poses 30000 --tag 30k_tile0 --save /tmp/base30k.npy >> $out 2>&1
ASAM_TILE_MODE=2 python tools/tune_batch.py --poses 30000 --tag 30k_tile2 --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
