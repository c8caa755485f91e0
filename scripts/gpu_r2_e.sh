#!/bin/bash
# Round-2 fifth GPU pass: DERIVE_MID A/B, large-FFT sweep with next-tile prefetch, ncu captures of the main launches.
set -u
mkdir -p gpurun_out
echo "== A/B DERIVE_MID"
: > gpurun_out/ab_dm.jsonl
for name in default dm1; do
  lib=pragma_dsp_b200/exp/lib_$name.so; [ $name = default ] && lib=pragma_dsp_b200/libpragma_b200.so
  [ -f $lib ] || continue
  timeout 300 python -u scripts/ab_tune.py --lib $lib --tag $name --key staged --values 0 --workloads north_star,c5,spectrum_f64,c3 --frames 1048576 --ms 120 >> gpurun_out/ab_dm.jsonl 2>> gpurun_out/ab_dm.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_dm.jsonl'):
    r = json.loads(l); print(r['tag'], r['workload'], '%.4f ms' % r['ms'], 'frac %.3f' % r['frac_of_measured_hbm'])
PY
tail -3 gpurun_out/ab_dm.err
echo "== C4 sweep"; python -u scripts/sweep_big.py > gpurun_out/sweep_big.txt 2> gpurun_out/sweep_big.err; cat gpurun_out/sweep_big.txt; tail -3 gpurun_out/sweep_big.err
for WL in north_star c2 c5 c3 c4_2e24; do
  K=r2c_kernel; case $WL in c4_*) K=bigfft;; esac
  CMD="python bench.py --workload $WL --steps 3 --warmup 3 --quick"
  $CMD > gpurun_out/plain_$WL.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$K -c 12 --csv \
      --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_list_$WL.log 2>&1
  echo "launch list $WL rc=$?"
  S=3; case $WL in c4_*) S=9;; esac
  $CMD > gpurun_out/plain2_$WL.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 \
      -o gpurun_out/prof_$WL -f $CMD > gpurun_out/ncu_full_$WL.log 2>&1
  echo "full capture $WL rc=$?"; tail -2 gpurun_out/ncu_full_$WL.log
done
# 2^24: the three passes of one transform
CMD="python bench.py --workload c4_2e24 --steps 3 --warmup 3 --quick"
$CMD > gpurun_out/plain3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bigfft -s 9 -c 3 -o gpurun_out/prof_c4_2e24 -f $CMD > gpurun_out/ncu_full_c4_2e24.log 2>&1
ls -la gpurun_out/*.ncu-rep
