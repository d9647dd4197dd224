#!/bin/bash
# second sweep: tiles per worker in team sizing (run under gpurun); $1 = log name
out=gpurun_out/${1:-tune2}.log
: > $out
python tools/tune_batch.py --tag tpw1 --save /tmp/base100k.npy >> $out 2>&1
for t in 2 3 4 6; do
  ASAM_TILES_PER_WORKER=$t timeout 120 python tools/tune_batch.py --tag tpw$t --check /tmp/base100k.npy >> $out 2>&1
done
ASAM_TILES_PER_WORKER=3 ASAM_TEAM_ROOM=148 timeout 120 python tools/tune_batch.py --tag tpw3_room148 --check /tmp/base100k.npy >> $out 2>&1
ASAM_TILES_PER_WORKER=3 ASAM_TEAM_ROOM=300 timeout 120 python tools/tune_batch.py --tag tpw3_room300 --check /tmp/base100k.npy >> $out 2>&1
python tools/tune_batch.py --poses 30000 --tag 30k_tpw1 --save /tmp/base30k.npy >> $out 2>&1
ASAM_TILES_PER_WORKER=3 python tools/tune_batch.py --poses 30000 --tag 30k_tpw3 --check /tmp/base30k.npy >> $out 2>&1
grep TUNE $out
