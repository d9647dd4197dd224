#!/bin/bash
# A/B runs of tuning knobs (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune}.log
: > $out
python tools/tune_batch.py --tag relax_2_24 --save /tmp/base100k.npy >> $out 2>&1
ASAM_RELAX_Z=4 ASAM_RELAX_FILL=64 timeout 120 python tools/tune_batch.py --tag relax_4_64 --check /tmp/base100k.npy >> $out 2>&1
ASAM_RELAX_Z=6 ASAM_RELAX_FILL=128 timeout 120 python tools/tune_batch.py --tag relax_6_128 --check /tmp/base100k.npy >> $out 2>&1
ASAM_RELAX_Z=8 ASAM_RELAX_FILL=256 timeout 120 python tools/tune_batch.py --tag relax_8_256 --check /tmp/base100k.npy >> $out 2>&1
python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_relax_2_24 --save /tmp/basem.npy >> $out 2>&1
ASAM_RELAX_Z=4 ASAM_RELAX_FILL=64 python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_relax_4_64 --check /tmp/basem.npy >> $out 2>&1
ASAM_RELAX_Z=6 ASAM_RELAX_FILL=128 python tools/tune_batch.py --workload m3500 --iters 30 --tag m3500_relax_6_128 --check /tmp/basem.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_relax_2_24 --save /tmp/base30k.npy >> $out 2>&1
ASAM_RELAX_Z=6 ASAM_RELAX_FILL=128 python tools/tune_batch.py --poses 30000 --tag 30k_relax_6_128 --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
