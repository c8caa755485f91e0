#!/bin/bash
# Round-2 second GPU pass: parity of everything new, kernel-configuration A/B over experiment builds, latency + ring lines.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -u -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then tail -20 gpurun_out/smoke.log; echo "smoke failed - stopping"; exit 1; fi
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== A/B experiment builds (north_star, c5, spectrum_f64)"
: > gpurun_out/ab_exp.jsonl
for name in default dl0 dp0 d00 mb5 mb6 mb5d00; do
  lib=pragma_dsp_b200/exp/lib_$name.so; [ $name = default ] && lib=pragma_dsp_b200/libpragma_b200.so
  [ -f $lib ] || continue
  timeout 300 python -u scripts/ab_tune.py --lib $lib --tag $name --key staged --values 0 --workloads north_star,c5,spectrum_f64 --frames 1048576 --ms 120 >> gpurun_out/ab_exp.jsonl 2>> gpurun_out/ab_exp.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_exp.jsonl'):
    r = json.loads(l); print(r['tag'], r['workload'], '%.4f ms' % r['ms'], 'frac %.3f' % r['frac_of_measured_hbm'])
PY
tail -3 gpurun_out/ab_exp.err
for wl in c1 north_star c2; do
  echo "== bench $wl"; timeout 400 python -u bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
