#!/bin/bash
# Output mode x size x precision sweep of the framed kernels (fraction of the measured HBM peak), one-sided
# specialised kernels and the two-sided generic kernel.  Writes gpurun_out/sweep_modes_sizes.jsonl.
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep_modes_sizes.jsonl
for m in amp amp_peak peak cplx amp_phase_peak; do
  SWEEP_MODE=$m timeout 300 python -u scripts/sweep_sizes.py 2>/dev/null >> gpurun_out/sweep_modes_sizes.jsonl
done
for m in amp amp_phase_peak; do
  SWEEP_SIDES=two SWEEP_MODE=$m SWEEP_LOG2N=6,8,10,12 timeout 300 python -u scripts/sweep_sizes.py 2>/dev/null >> gpurun_out/sweep_modes_sizes.jsonl
done
wc -l gpurun_out/sweep_modes_sizes.jsonl
