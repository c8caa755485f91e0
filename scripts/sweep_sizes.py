#!/usr/bin/env python3
"""Fraction of the HBM roofline across FFT sizes (amplitude + peak, Hann) for both precisions.

    python scripts/sweep_sizes.py > gpurun_out/sweep_sizes.jsonl
    SWEEP_MODE=amp|amp_peak|peak|cplx|amp_phase_peak SWEEP_SIDES=one|two SWEEP_LOG2N=10,11,12 python scripts/sweep_sizes.py
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pragma_dsp_b200 import _lib  # noqa: E402
from pragma_dsp_b200._lib import F32, F64, SIDES, WINDOWS, SpectrumDesc, check, lib  # noqa: E402

ctx = _lib.Context(0)
L = lib()
peak, _ = bench.measured_hbm_peak()
MODE = os.environ.get("SWEEP_MODE", "amp_peak")
SIDES_ = os.environ.get("SWEEP_SIDES", "one")
PAD = int(os.environ.get("SWEEP_PAD", "0"))  # frames are PAD samples shorter than N (zero-padded by the kernel)
SIZES = [int(v) for v in os.environ.get("SWEEP_LOG2N", ",".join(str(i) for i in range(5, 15))).split(",")]
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
for prec_name, prec, tdt in (("f64", F64, torch.float64), ("f32", F32, torch.float32)):
    for log2n in SIZES:
        n = 1 << log2n
        frames = max(64, (1 << 26) // n)  # ~64M samples per launch: far beyond L2
        flen = max(2, n - PAD) & ~1
        x = torch.randn((frames, flen), device=dev, dtype=tdt)
        bins = n // 2 + 1 if SIDES_ == "one" else n
        amp = torch.empty((frames, bins), dtype=tdt, device=dev)
        pk = torch.zeros((frames, 32), dtype=torch.uint8, device=dev)
        plan = ctx.plan(n, prec)
        d = SpectrumDesc(sample_dtype=prec, frame_len=flen, hop=flen, batch=frames, window=WINDOWS["hann"], sides=SIDES[SIDES_],
                         sample_rate=48000.0, raw_magnitude=0)

        if MODE == "cplx":  # Radix2Fft.forward: all N bins, two planes
            amp = torch.empty((frames, n), dtype=tdt, device=dev)
        im = torch.empty((frames, n), dtype=tdt, device=dev) if MODE == "cplx" else None
        ph = torch.empty((frames, bins), dtype=tdt, device=dev) if "phase" in MODE else None

        def go():
            if MODE == "cplx":
                check(L.pdsp_fft_forward_real_dev(plan, C.c_void_p(x.data_ptr()), prec, frames, C.c_void_p(amp.data_ptr()),
                                                  C.c_void_p(im.data_ptr()), 1, C.c_void_p(st.cuda_stream)))
                return
            check(L.pdsp_spectrum_dev(plan, C.byref(d), C.c_void_p(x.data_ptr()),
                                      C.c_void_p(amp.data_ptr()) if "amp" in MODE else None,
                                      C.c_void_p(ph.data_ptr()) if ph is not None else None,
                                      C.c_void_p(pk.data_ptr()) if "peak" in MODE else None, C.c_void_p(st.cuda_stream)))
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            go()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        es = 8 if prec == F64 else 4
        bpf = flen * es + (2 * n * es if MODE == "cplx" else (bins * es if "amp" in MODE else 0) + (bins * es if "phase" in MODE else 0)
                        + ((32 if prec == F64 else 16) if "peak" in MODE else 0))
        gbs = frames * bpf / (ms * 1e-3) / 1e9
        print(json.dumps({"mode": MODE, "sides": SIDES_, "pad": PAD, "precision": prec_name, "n": n, "frames": frames, "ms": ms, "frames_per_s": frames / (ms * 1e-3),
                          "gbs": gbs, "frac_of_measured_hbm": gbs / peak}), flush=True)
        del x, amp, pk, im, ph
