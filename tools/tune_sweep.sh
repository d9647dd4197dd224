#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
python tools/tune_batch.py --tag base --save /tmp/base100k.npy >> $out 2>&1
for pc in 20 30 40 60; do
ASAM_TEAM_MERGE_PCT=$pc timeout 120 python tools/tune_batch.py --tag team_merge_$pc --check /tmp/base100k.npy >> $out 2>&1
done
ASAM_TEAM_MERGE_PCT=30 ASAM_TEAM_MERGE_MFLOP=150 timeout 120 python tools/tune_batch.py --tag team_merge_30_150 --check /tmp/base100k.npy >> $out 2>&1
ASAM_TEAM_MERGE_PCT=40 ASAM_TEAM_MERGE_MFLOP=1000 timeout 120 python tools/tune_batch.py --tag team_merge_40_1000 --check /tmp/base100k.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_base --save /tmp/basem.npy >> $out 2>&1
ASAM_TEAM_MERGE_PCT=30 python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_team_merge_30 --check /tmp/basem.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_base --save /tmp/base30k.npy >> $out 2>&1
ASAM_TEAM_MERGE_PCT=30 python tools/tune_batch.py --poses 30000 --tag 30k_team_merge_30 --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
