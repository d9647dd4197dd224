#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_SMEM_MMA=0 python tools/tune_batch.py --tag dfma_smem --save /tmp/base100k.npy >> $out 2>&1
timeout 120 python tools/tune_batch.py --tag mma_smem --check /tmp/base100k.npy >> $out 2>&1
ASAM_SMEM_MMA=0 python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_dfma_smem --save /tmp/basem.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_mma_smem --check /tmp/basem.npy >> $out 2>&1
ASAM_SMEM_MMA=0 python tools/tune_batch.py --poses 30000 --tag 30k_dfma_smem --save /tmp/base30k.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_mma_smem --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
