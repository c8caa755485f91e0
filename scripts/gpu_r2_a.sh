#!/bin/bash
# Round-2 first GPU pass: smoke, parity tests, staged A/B, bench lines of every workload, reference arm.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -u -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed - stopping"; exit 1; fi
echo "== A/B staged"; timeout 600 python -u scripts/ab_tune.py --key staged --values 0,1 > gpurun_out/ab_staged.jsonl 2> gpurun_out/ab_staged.err; echo "ab rc=$?"; cat gpurun_out/ab_staged.jsonl | cut -c1-260; tail -3 gpurun_out/ab_staged.err
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
for wl in north_star c2 c5 spectrum_f64 c3 c4_2e20 c4_2e24 c1; do
  echo "== bench $wl"; timeout 400 python -u bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
echo "== reference arm"; timeout 400 python -u bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_reference.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.csv 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; free -g >> gpurun_out/lscpu.txt
