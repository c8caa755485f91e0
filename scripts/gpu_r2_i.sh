#!/bin/bash
# Round-2 ninth GPU pass: ncu capture of the third-generation (pipeline) passes of 2^20 (both passes of one step).
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload c4_2e20 --steps 3 --warmup 3 --quick"
$CMD > gpurun_out/plain_pipe_c4_2e20.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/plain_pipe_c4_2e20.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bigfft -c 12 --csv --log-file gpurun_out/launches_c4_2e20.csv $CMD > gpurun_out/ncu_list_c4_2e20.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bigfft -s 6 -c 2 -o gpurun_out/prof_c4_2e20 -f $CMD > gpurun_out/ncu_full_c4_2e20.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full_c4_2e20.log
