#!/bin/bash
# Round-2 eighth GPU pass: third-generation (pipeline) large-FFT passes - parity, sweep against the earlier generations.
set -u
mkdir -p gpurun_out
echo "== large-FFT parity tests"; timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "large or multipass or big or c4" > gpurun_out/pytest_gpu_big.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_big.log
echo "== pipe sweep"; timeout 900 python -u scripts/sweep_big.py --pipe > gpurun_out/sweep_big_pipe.txt 2> gpurun_out/sweep_big_pipe.err; cat gpurun_out/sweep_big_pipe.txt; tail -3 gpurun_out/sweep_big_pipe.err
