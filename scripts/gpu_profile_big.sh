#!/bin/bash
# K2 first generation: ncu full captures of the 2^24 passes with per-thread (PDSP_BIG_TMA=0) and TMA (=1) tile loads.
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload c4_2e24 --steps 2 --warmup 3 --quick"
PDSP_BIG_TMA=0 $CMD > gpurun_out/plain_c4.log 2>&1 &&
PDSP_BIG_TMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:bigfft -s 9 -c 3 -o gpurun_out/prof_c4_plain -f $CMD > gpurun_out/ncu_c4_plain.log 2>&1
echo "plain rc=$?"
PDSP_BIG_TMA=1 $CMD > gpurun_out/plain_c4t.log 2>&1 &&
PDSP_BIG_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:bigfft -s 9 -c 3 -o gpurun_out/prof_c4_tma -f $CMD > gpurun_out/ncu_c4_tma.log 2>&1
echo "tma rc=$?"
