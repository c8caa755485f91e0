#!/bin/bash
# Pass-length factorisations of the large FFT (PDSP_BIG_FACTORS = log2 of each pass length), timing + parity per line.
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep_big_factors.txt
run() {
  PDSP_BIG_FACTORS=$2 timeout -s KILL 120 python bench.py --workload $1 --steps 30 --warmup 5 --quick 2>/dev/null | python -c "
import sys, json
try:
    r = json.loads(sys.stdin.read()); print('factors=$2', r['config']['workload'], '%.4f ms/step' % r['ms_per_step'], 'frac %.3f' % r['roofline']['frac'], 'rel_l2 %.2e' % r['parity']['rel_l2_vs_numpy'])
except Exception as e:
    print('factors=$2 $1 failed', e)
" >> gpurun_out/sweep_big_factors.txt
}
for f in 10,10 7,7,6 6,7,7 8,6,6 6,6,8 7,6,7; do run c4_2e20 $f; done
for f in 8,8,8 9,9,6 6,9,9 10,7,7 7,7,10 9,8,7 7,8,9 8,9,7; do run c4_2e24 $f; done
cat gpurun_out/sweep_big_factors.txt
