#!/bin/bash
# Round-2 third GPU pass: full GPU test suite; large-FFT second generation vs first, chunk sweep.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== C4 sweeps"
python -u scripts/sweep_big.py > gpurun_out/sweep_big.txt 2> gpurun_out/sweep_big.err; echo "rc=$?"; cat gpurun_out/sweep_big.txt; tail -5 gpurun_out/sweep_big.err
for wl in c4_2e20 c4_2e24; do
  echo "== bench $wl"; timeout 400 python -u bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
