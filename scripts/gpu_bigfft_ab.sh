#!/bin/bash
# A/B of the large-N passes: TMA vs plain tile loads x L2 prefetch of the next tile. Parity is part of each bench line.
set -u
mkdir -p gpurun_out
for il in 0 1; do for pf in 0 1; do
  for mode in 1 0; do
    for wl in c4_2e20 c4_2e24; do
      PDSP_BIG_INTERLEAVE=$il PDSP_BIG_PREFETCH=$pf PDSP_BIG_TMA=$mode timeout -s KILL 200 python -u bench.py --workload $wl --steps 20 --warmup 3 --quick > gpurun_out/bench_${wl}_tma${mode}_pf$pf.json 2> gpurun_out/bench_${wl}_tma${mode}_pf$pf.err; python - <<PY
import json
try:
    r = json.load(open("gpurun_out/bench_${wl}_tma${mode}_pf$pf.json")); print("$wl interleave=$il tma=$mode prefetch=$pf", "%.4f ms/step" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], r["parity"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${wl}_tma${mode}_pf$pf.err").read()[-600:])
PY
    done
  done
done
done
