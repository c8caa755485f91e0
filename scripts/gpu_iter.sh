#!/bin/bash
# Iteration run: smoke, parity tests, variant sweep, optional extra bench workloads.
# Usage: bash scripts/gpu_iter.sh ["workloads for bench"]
set -u
WLS=${1:-""}
mkdir -p gpurun_out
echo "== smoke"; timeout 240 python -u -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed - stopping"; exit 1; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== variant sweep"; timeout 300 python -u scripts/sweep_variants.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    r = json.loads(l); print(r['workload'], 'v%d' % r['variant'], '%.3f ms' % r['ms'], '%.3e f/s' % r['frames_per_s'], '%.0f GB/s' % r['gbs'], '%.3f' % r['frac_of_measured_hbm'], r['agree_with_v0'])
PY
tail -3 gpurun_out/sweep.err
if [ -n "${AB_ENV:-}" ]; then
  echo "== A/B: $AB_ENV"; env $AB_ENV timeout 300 python -u scripts/sweep_variants.py --variants 0 > gpurun_out/sweep_ab.jsonl 2>> gpurun_out/sweep.err; python - <<'PY'
import json
for l in open('gpurun_out/sweep_ab.jsonl'):
    r = json.loads(l); print('AB', r['workload'], 'v%d' % r['variant'], '%.3f ms' % r['ms'], '%.3e f/s' % r['frames_per_s'], '%.3f' % r['frac_of_measured_hbm'])
PY
fi
for wl in $WLS; do
  echo "== bench $wl"; timeout 300 python -u bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench rc=$?"; cat gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
