#!/usr/bin/env python3
"""End-to-end (pinned host buffers -> pdsp_spectrum -> pinned host buffers) throughput of the C2 workload as a function
of the staging chunk size (PDSP_CHUNK_BYTES test hook).  python scripts/e2e_chunk_sweep.py"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pragma_dsp_b200 import _lib  # noqa: E402
from pragma_dsp_b200._lib import F32, SIDES, WINDOWS, SpectrumDesc, check, lib  # noqa: E402

ctx = _lib.Context(0)
L = lib()
n, frames = 1024, 65536
bins = n // 2 + 1
hx = torch.randn((frames, n), dtype=torch.float32).pin_memory()
h_amp = torch.empty((frames, bins), dtype=torch.float32).pin_memory()
h_pk = torch.empty((frames, 16), dtype=torch.uint8).pin_memory()
plan = ctx.plan(n, F32)
d = SpectrumDesc(sample_dtype=F32, frame_len=n, hop=n, batch=frames, window=WINDOWS["hann"], sides=SIDES["one"],
                 sample_rate=48000.0, raw_magnitude=0, fft_shift=0)


def step():
    check(L.pdsp_spectrum(plan, C.byref(d), C.c_void_p(hx.data_ptr()), C.c_void_p(h_amp.data_ptr()), None,
                          C.c_void_p(h_pk.data_ptr())))


for mb in (4, 8, 16, 24, 32, 48, 64, 96, 128):
    ctx.tune("chunk_bytes", mb << 20)
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    print(f"chunk {mb:4d} MB: {frames / dt:.3e} frames/s  H2D {hx.numel() * 4 / dt / 1e9:.1f} GB/s  D2H {(h_amp.numel() * 4 + h_pk.numel()) / dt / 1e9:.1f} GB/s", flush=True)
