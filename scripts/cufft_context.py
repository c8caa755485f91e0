#!/usr/bin/env python3
"""CONTEXT ONLY (SURVEY 8d "optional sanity comparator"): what a stock library reaches on the same box for the same frames -
torch.fft.rfft (cuFFT D2Z / R2C) alone, and the unfused pipeline window * x -> rfft -> abs.  Not a product path."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

peak, _ = bench.measured_hbm_peak()
dev = torch.device("cuda", 0)
for n, dt, frames in ((1024, torch.float64, 1 << 19), (1024, torch.float32, 1 << 20), (4096, torch.float64, 1 << 17)):
    es = 8 if dt == torch.float64 else 4
    x = torch.randn((frames, n), device=dev, dtype=dt)
    w = torch.hann_window(n, periodic=False, device=dev, dtype=dt)

    def timed(fn, ms=150.0):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        reps = max(3, int(ms / max(e0.elapsed_time(e1), 1e-3)))
        for _ in range(reps):  # warm-up of the same length: sustained clocks
            fn()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_fft = timed(lambda: torch.fft.rfft(x, dim=1))
    t_pipe = timed(lambda: torch.fft.rfft(x * w, dim=1).abs())
    bins = n // 2 + 1
    for what, t, bpf in (("rfft only (in N, out N/2+1 complex)", t_fft, n * es + bins * 2 * es),
                         ("x*window -> rfft -> abs (algorithmic bytes: in N, out N/2+1)", t_pipe, n * es + bins * es)):
        fps = frames / (t * 1e-3)
        print(json.dumps({"n": n, "dtype": str(dt), "frames": frames, "what": what, "ms": t, "frames_per_s": fps,
                          "frac_of_measured_hbm_on_algorithmic_bytes": fps * bpf / 1e9 / peak}), flush=True)
    del x

# large single complex transforms (config C4): cuFFT Z2Z through torch.fft.fft, 32 bytes per point algorithmic
for log2n, nf in ((16, 64), (18, 16), (20, 8), (22, 2), (24, 1), (26, 1)):
    n = 1 << log2n
    z = torch.complex(torch.rand((nf, n), device=dev, dtype=torch.float64) * 2 - 1, torch.rand((nf, n), device=dev, dtype=torch.float64) * 2 - 1)

    def timed2(fn, ms=100.0):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        reps = max(5, int(ms / max(e0.elapsed_time(e1), 1e-3)))
        for _ in range(reps):
            fn()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t = timed2(lambda: torch.fft.fft(z, dim=1))
    print(json.dumps({"n": n, "log2n": log2n, "dtype": "complex128", "transforms": nf, "what": "torch.fft.fft (cuFFT Z2Z, interleaved in/out)",
                      "ms": t, "us_per_transform": t / nf * 1e3, "frac_of_measured_hbm_on_algorithmic_bytes": 32.0 * n * nf / (t * 1e-3) / 1e9 / peak}), flush=True)
    del z
