#!/bin/bash
# Variant sweep + ncu launch list + full capture of the dominant kernel for the given workloads.
# Usage: bash scripts/gpu_profile.sh "c2 north_star" ; ncu only runs after the plain command exited 0.
set -u
WLS=${1:-"c2 north_star"}
mkdir -p gpurun_out
echo "== variant sweep"; timeout 300 python -u scripts/sweep_variants.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    r = json.loads(l); print(r['workload'], 'v%d' % r['variant'], '%.3f ms' % r['ms'], '%.3e f/s' % r['frames_per_s'], '%.0f GB/s' % r['gbs'], '%.3f' % r['frac_of_measured_hbm'], r['agree_with_v0'])
PY
tail -3 gpurun_out/sweep.err
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
