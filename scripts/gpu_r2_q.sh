#!/bin/bash
# A/B: linearised exchange gather (now the default build) and fp64 root constants from the constant bank (ct64).
set -u
mkdir -p gpurun_out
: > gpurun_out/ab_linear_gather.jsonl
for name in default ct64 default ct64; do
  lib=pragma_dsp_b200/exp/lib_$name.so; [ $name = default ] && lib=pragma_dsp_b200/libpragma_b200.so
  [ -f $lib ] || continue
  timeout 400 python -u scripts/ab_tune.py --lib $lib --tag $name --key staged --values 0 --workloads north_star,c5,c2,spectrum_f64,c3 --frames 1048576 --ms 120 >> gpurun_out/ab_linear_gather.jsonl 2>> gpurun_out/ab_linear_gather.err
done
python - <<'PY'
import json, collections
t = collections.OrderedDict()
for l in open('gpurun_out/ab_linear_gather.jsonl'):
    r = json.loads(l); t.setdefault(r['tag'], []).append((r['workload'], r['frac_of_measured_hbm']))
for k, v in t.items():
    print('%-8s' % k, ' '.join('%s %.3f' % (w, f) for w, f in v))
PY
tail -3 gpurun_out/ab_linear_gather.err
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "forward_real_all_sizes or spectrum_batch_vs_oracle or large_multipass or replays_reference" 2>&1 | tail -3
