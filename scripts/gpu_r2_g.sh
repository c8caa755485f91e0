#!/bin/bash
# Round-2 seventh GPU pass: parity of the grouped (L2-resident) three-pass schedule, its sweep, full GPU test suite.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 240 python -u -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== resident sweep"; timeout 600 python -u scripts/sweep_big.py --resident > gpurun_out/sweep_big_resident.txt 2> gpurun_out/sweep_big_resident.err; cat gpurun_out/sweep_big_resident.txt; tail -3 gpurun_out/sweep_big_resident.err
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
