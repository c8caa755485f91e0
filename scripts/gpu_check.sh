#!/bin/bash
# First-contact GPU run: smoke, parity tests, bench lines, ncu launch list + one full capture.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for wl in c2 north_star c5; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench rc=$?"; cat gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
