#!/bin/bash
# Round-2 sixth GPU pass: staged-load north-star capture, C1 latency and C2 bench lines.
set -u
mkdir -p gpurun_out
export PDSP_STAGED=1
for WL in north_star; do
  CMD="python bench.py --workload $WL --steps 3 --warmup 3 --quick"
  $CMD > gpurun_out/plain_staged_$WL.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:r2c_kernel -s 3 -c 1 \
      -o gpurun_out/prof_staged_$WL -f $CMD > gpurun_out/ncu_full_staged_$WL.log 2>&1
  echo "full capture staged $WL rc=$?"; tail -2 gpurun_out/ncu_full_staged_$WL.log
done
unset PDSP_STAGED
echo "== bench c1"; timeout 400 python -u bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "rc=$?"; cut -c1-900 gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
echo "== bench c2"; timeout 400 python -u bench.py --workload c2 --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
