#!/bin/bash
# Round-2 GPU pass: A/B of the next-frame prefetch hint (CCTL.E.PF1 / PF2) in the framed kernels.
set -u
mkdir -p gpurun_out
: > gpurun_out/ab_prefetch.jsonl
for name in default pf1 pf2 default; do
  lib=pragma_dsp_b200/exp/lib_$name.so; [ $name = default ] && lib=pragma_dsp_b200/libpragma_b200.so
  [ -f $lib ] || continue
  timeout 400 python -u scripts/ab_tune.py --lib $lib --tag $name --key staged --values 0 --workloads north_star,c5,c2,spectrum_f64,c3 --frames 1048576 --ms 120 >> gpurun_out/ab_prefetch.jsonl 2>> gpurun_out/ab_prefetch.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_prefetch.jsonl'):
    r = json.loads(l); print(r['tag'], r['workload'], '%.4f ms' % r['ms'], 'frac %.3f' % r['frac_of_measured_hbm'])
PY
tail -3 gpurun_out/ab_prefetch.err
