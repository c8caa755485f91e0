#!/bin/bash
# ncu launch list + full capture of the dominant kernel for the given workloads.
# Usage: bash scripts/gpu_profile.sh "c2 north_star" ; ncu only runs after the plain command exited 0.
set -u
WLS=${1:-"c2 north_star"}
mkdir -p gpurun_out
for WL in $WLS; do
  CMD="python bench.py --workload $WL --steps 3 --warmup 3 --quick"
  $CMD > gpurun_out/plain_$WL.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:r2c_kernel -c 40 --csv \
      --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_list_$WL.log 2>&1
  echo "launch list $WL rc=$?"
  $CMD > gpurun_out/plain2_$WL.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:r2c_kernel -s 4 -c 1 \
      -o gpurun_out/prof_$WL -f $CMD > gpurun_out/ncu_full_$WL.log 2>&1
  echo "full capture $WL rc=$?"; tail -2 gpurun_out/ncu_full_$WL.log
done
ls -la gpurun_out/
