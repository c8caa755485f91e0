#!/usr/bin/env python3
"""Small, fast exercise of every kernel family for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool racecheck python scripts/sanitize_target.py
    compute-sanitizer --tool memcheck  python scripts/sanitize_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from pragma_dsp_b200 import spectrum_batch  # noqa: E402
from pragma_dsp_b200.core import ComplexArray, Radix2Fft  # noqa: E402

rng = np.random.default_rng(0)
# r2c: generic + specialised kernels, warp-per-frame, multi-warp frames, tiny sizes, both precisions
for n, batch in ((2, 5), (8, 7), (64, 9), (256, 33), (1024, 67), (4096, 9), (16384, 2)):
    x = rng.standard_normal((batch, n))
    for prec in ("f64", "f32"):
        xs = x.astype(np.float32) if prec == "f32" else x
        for outs in (("amplitude", "phase", "peak"), ("amplitude", "peak"), ("peak",), ("amplitude",)):
            spectrum_batch(xs, sampleRate=48000.0, fftSize=n, window="hann", precision=prec, outputs=outs)
    spectrum_batch(x[:, : n // 2 + 1], sampleRate=1.0, fftSize=n, window="rect", sides="two")  # generic, zero-padded
    f = Radix2Fft(n)
    re, im = f.forward_batch(x)
    f.complex_batch(re, im, inverse=True)
# multi-pass large transforms, both load paths
for tma in ("1", "0"):
    os.environ["PDSP_BIG_TMA"] = tma
    for log2n in (14, 16, 21):
        n = 1 << log2n
        c = ComplexArray(rng.standard_normal(n), rng.standard_normal(n))
        out = Radix2Fft(n).forwardComplex(c)
        ref = np.fft.fft(c.real + 1j * c.imag)
        assert np.linalg.norm(out.real + 1j * out.imag - ref) / np.linalg.norm(ref) < 1e-11
print("sanitize target ok")
