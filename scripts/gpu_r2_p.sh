#!/bin/bash
# A/B of cache-policy hints: sample loads (.cs / L1::no_allocate / .lu), streaming amplitude stores, evict-last table loads.
set -u
mkdir -p gpurun_out
: > gpurun_out/ab_cache_policy.jsonl
for name in default ld1 ld2 ld3 ld2st tl1 ld2tl default; do
  lib=pragma_dsp_b200/exp/lib_$name.so; [ $name = default ] && lib=pragma_dsp_b200/libpragma_b200.so
  [ -f $lib ] || continue
  timeout 400 python -u scripts/ab_tune.py --lib $lib --tag $name --key staged --values 0 --workloads north_star,c5,c2,spectrum_f64,c3 --frames 1048576 --ms 120 >> gpurun_out/ab_cache_policy.jsonl 2>> gpurun_out/ab_cache_policy.err
done
python - <<'PY'
import json, collections
t = collections.OrderedDict()
for l in open('gpurun_out/ab_cache_policy.jsonl'):
    r = json.loads(l); t.setdefault(r['tag'], []).append((r['workload'], r['frac_of_measured_hbm']))
for k, v in t.items():
    print('%-8s' % k, ' '.join('%s %.3f' % (w, f) for w, f in v))
PY
tail -3 gpurun_out/ab_cache_policy.err
