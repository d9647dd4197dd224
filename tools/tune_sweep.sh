#!/bin/bash
# tuning sweep of the single-CTA ("solo") path for mid-size fronts (run under gpurun)
out=gpurun_out/${1:-tune}.log
: > $out
python tools/tune_batch.py --tag base --save /tmp/base100k.npy >> $out 2>&1
for pb in 48 24; do
  for mm in 240 330 450 600 900; do
    ASAM_SOLO_MAX_M=$mm ASAM_SOLO_PB=$pb python tools/tune_batch.py --tag solo${mm}_pb${pb} --check /tmp/base100k.npy >> $out 2>&1
  done
done
python tools/tune_batch.py --workload m3500 --tag m3500_base --save /tmp/basem.npy >> $out 2>&1
grep TUNE $out
