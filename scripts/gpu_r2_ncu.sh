#!/bin/bash
# Refresh the ncu evidence of the framed kernels with the final code: launch list + one full capture per workload.
set -u
mkdir -p gpurun_out
for WL in north_star c2 c5 c3; do
  CMD="python bench.py --workload $WL --steps 3 --warmup 3 --quick"
  $CMD > gpurun_out/plain_$WL.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:r2c_kernel -c 12 --csv \
      --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_list_$WL.log 2>&1
  echo "launch list $WL rc=$?"
  $CMD > gpurun_out/plain2_$WL.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:r2c_kernel -s 3 -c 1 \
      -o gpurun_out/prof_$WL -f $CMD > gpurun_out/ncu_full_$WL.log 2>&1
  echo "full capture $WL rc=$?"; tail -1 gpurun_out/ncu_full_$WL.log
done
ls -la gpurun_out/prof_north_star.ncu-rep gpurun_out/prof_c2.ncu-rep gpurun_out/prof_c5.ncu-rep gpurun_out/prof_c3.ncu-rep
