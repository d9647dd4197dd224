#!/bin/bash
# Evidence of a round (run under gpurun, ONE GPU): GPU tests, smoke, the default bench line (all four workloads) of both
# arms, ncu launch lists + --set full captures, a compute-sanitizer pass.  $1 = file prefix (default rd2).
R=${1:-rd2}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/${R}_pytest_tail.log
cat gpurun_out/${R}_pytest_tail.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_n1_reference.json 2>> gpurun_out/${R}.err
tail -3 gpurun_out/${R}.err
# the sparse-world replay over two more windows (the default line times steps 50000..): early and latest sound S
timeout 600 python bench.py --workload manhattan_replay --replay-from 10000 --no-cpu-baseline > gpurun_out/${R}_replay_from10k.json 2>> gpurun_out/${R}.err
timeout 600 python bench.py --workload manhattan_replay --replay-from 90000 --no-cpu-baseline > gpurun_out/${R}_replay_from90k.json 2>> gpurun_out/${R}.err
timeout 300 python tools/step_profile.py > gpurun_out/${R}_step_profile.log 2>&1
timeout 300 python tools/panel_trace.py --fronts 2 > gpurun_out/${R}_panel_trace.log 2>&1
bash tools/profile_round.sh ${R} full 2>&1 | tail -12
timeout 900 compute-sanitizer --tool memcheck python tools/gpu_diag.py --sizes 100,3500 --replay 60 --synth 12000 > gpurun_out/${R}_memcheck.log 2>&1
tail -3 gpurun_out/${R}_memcheck.log
