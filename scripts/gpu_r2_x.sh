#!/bin/bash
# Round-2 GPU pass: large-FFT parity with the final per-pass selection, C4 bench lines, launch list + full capture of the 2^24 passes.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "large or multipass or generations or c4" 2>&1 | tail -3
for WL in c4_2e24 c4_2e20; do
  timeout 300 python -u bench.py --workload $WL > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "bench $WL rc=$?"; cut -c1-200 gpurun_out/bench_$WL.json
done
CMD="python bench.py --workload c4_2e24 --steps 3 --warmup 3 --quick"
$CMD > gpurun_out/plain_c4_2e24.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bigfft -c 12 --csv --log-file gpurun_out/launches_c4_2e24.csv $CMD > gpurun_out/ncu_list_c4_2e24.log 2>&1; echo "list rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bigfft -s 9 -c 3 -o gpurun_out/prof_c4_2e24 -f $CMD > gpurun_out/ncu_full_c4_2e24.log 2>&1; echo "full rc=$?"
