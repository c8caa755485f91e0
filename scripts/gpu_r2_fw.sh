#!/bin/bash
# A/B of the window multiply fused into the first butterfly stage (-DPDSP_FUSE_WINDOW=1) + a parity check of that build.
set -u
mkdir -p gpurun_out
: > gpurun_out/ab_fuse_window.jsonl
for name in default fw default fw; do
  lib=pragma_dsp_b200/exp/lib_$name.so; [ $name = default ] && lib=pragma_dsp_b200/libpragma_b200.so
  [ -f $lib ] || continue
  timeout 400 python -u scripts/ab_tune.py --lib $lib --tag $name --key staged --values 0 --workloads north_star,c5,spectrum_f64,c3 --frames 1048576 --ms 150 >> gpurun_out/ab_fuse_window.jsonl 2>> gpurun_out/ab_fuse_window.err
done
python - <<'PY'
import json, collections
t = collections.OrderedDict()
for l in open('gpurun_out/ab_fuse_window.jsonl'):
    r = json.loads(l); t.setdefault(r['tag'], []).append((r['workload'], r['frac_of_measured_hbm']))
for k, v in t.items():
    print('%-8s' % k, ' '.join('%s %.4f' % (w, f) for w, f in v))
PY
tail -3 gpurun_out/ab_fuse_window.err
python - <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from pragma_dsp_b200 import _lib
_lib.LIB_PATH = os.path.abspath("pragma_dsp_b200/exp/lib_fw.so")
import oracle
from pragma_dsp_b200 import spectrum_batch
rng = np.random.default_rng(5)
worst = 0.0
for n in (8, 64, 256, 1024, 2048, 4096, 16384):
    for window in ("rect", "hann", "blackman"):
        x = rng.standard_normal((37, n)) + 3 * np.sin(2 * np.pi * 5 * np.arange(n) / n)
        got = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window=window, precision="f64")
        ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window=window)
        err = np.abs(got["amplitude"] - ref["amplitude"]).max()
        worst = max(worst, err)
        assert err <= 1e-12, (n, window, err)
        assert (got["peaks"]["index"] == ref["peaks"]["index"]).all(), (n, window)
    z = spectrum_batch(np.zeros((3, n)), fftSize=n, window="hann", precision="f64")
    assert (z["amplitude"] == 0).all()
    c = spectrum_batch(np.full((3, n), 0.75), fftSize=n, window="rect", precision="f64")
    assert (c["amplitude"][:, 1:] == 0).all() and (c["peaks"]["index"] == 0).all(), n
print("fused-window build parity ok, worst amplitude error", worst)
PY
