#!/bin/bash
# Large-N (multi-pass) FFT check: TMA-staged passes vs plain, parity first under a hard timeout, then timing.
set -u
mkdir -p gpurun_out
cat > /tmp/big_check.py <<'PY'
import sys, time
import numpy as np
sys.path.insert(0, '.')
from pragma_dsp_b200.core import Radix2Fft, ComplexArray
for log2n in (14, 17, 20, 21, 24):
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    re, im = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    out = Radix2Fft(n).forwardComplex(ComplexArray(re, im))
    ref = np.fft.fft(re + 1j * im)
    err = np.linalg.norm(out.real + 1j * out.imag - ref) / np.linalg.norm(ref)
    print(f"2^{log2n}: rel-L2 {err:.2e} (bound {1e-12 * log2n:.1e})", flush=True)
    assert err <= 1e-12 * log2n
print("big fft parity ok")
PY
for mode in 1 0; do
  echo "== PDSP_BIG_TMA=$mode parity"; PDSP_BIG_TMA=$mode timeout -s KILL 120 python -u /tmp/big_check.py 2>&1 | tail -7; echo "rc=$?"
done
for mode in 1 0; do
  for wl in c4_2e20 c4_2e24; do
    echo "== bench $wl PDSP_BIG_TMA=$mode"; PDSP_BIG_TMA=$mode timeout -s KILL 200 python -u bench.py --workload $wl --steps 20 --warmup 3 --quick > gpurun_out/bench_${wl}_tma$mode.json 2> gpurun_out/bench_${wl}_tma$mode.err; python - <<PY
import json
try:
    r = json.load(open("gpurun_out/bench_${wl}_tma$mode.json")); print("$wl tma=$mode", "%.4f ms/step" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], r["parity"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${wl}_tma$mode.err").read()[-600:])
PY
  done
done
