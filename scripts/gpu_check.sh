#!/bin/bash
# GPU session: smoke, parity tests, bench lines, variant sweep.  Every stage has its own timeout.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== smoke"; timeout 240 python -u -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -5 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed - stopping"; exit 1; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
for wl in c2 north_star c5 c3 c4_2e20 c4_2e24 run_ts_2048 run_ts_4096; do
  echo "== bench $wl"; timeout 300 python -u bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench rc=$?"; cat gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
echo "== reference arm"; timeout 300 python -u bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
