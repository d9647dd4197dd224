#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
ASAM_STAGED=1 ASAM_DIAG_MMA=0 python tools/tune_batch.py --tag base --save /tmp/base100k.npy >> $out 2>&1
ASAM_DIAG_MMA=0 timeout 120 python tools/tune_batch.py --tag crew_pipelined --check /tmp/base100k.npy >> $out 2>&1
ASAM_STAGED=1 timeout 120 python tools/tune_batch.py --tag diag_mma --check /tmp/base100k.npy >> $out 2>&1
timeout 120 python tools/tune_batch.py --tag both --check /tmp/base100k.npy >> $out 2>&1
ASAM_STAGED=1 ASAM_DIAG_MMA=0 python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_base --save /tmp/basem.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_both --check /tmp/basem.npy >> $out 2>&1
ASAM_STAGED=1 ASAM_DIAG_MMA=0 python tools/tune_batch.py --poses 30000 --tag 30k_base --save /tmp/base30k.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_both --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
