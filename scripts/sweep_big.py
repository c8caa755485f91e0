#!/usr/bin/env python3
"""Large-FFT (K2) sweep on the GPU: generation (big_v2), transforms per group of passes (big_chunk), work-buffer layout,
for N = 2^20 (8 transforms) and 2^24 (1).  Prints ms per step and the fraction of the HBM roofline (32 B per point)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from pragma_dsp_b200 import _lib  # noqa: E402

if os.environ.get("PDSP_LIB"):  # time another build of the library (scripts/build_exp.sh)
    _lib.LIB_PATH = os.path.abspath(os.environ["PDSP_LIB"])
from pragma_dsp_b200._lib import F64, check, lib  # noqa: E402

ctx = _lib.Context(0)
L = lib()
peak, _ = bench.measured_hbm_peak()
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
KEYS = ("big_v2", "big_chunk", "big_interleave", "big_resident", "big_pipe", "big_factors", "big_fused")
cases = [(20, 8), (24, 1), (22, 2), (16, 64), (18, 16)]
settings = [dict(), dict(big_v2=0), dict(big_v2=1), dict(big_v2=0, big_chunk=2), dict(big_v2=0, big_chunk=4)]
if "--resident" in sys.argv:  # 2^24 / 2^26: passes 1+2 in L2-sized k1 groups (big_resident = blocks per group), against pass-by-pass
    cases = [(24, 1), (26, 1)]
    settings = [dict(big_resident=0), dict(), dict(big_resident=16), dict(big_resident=32), dict(big_resident=64)]
if "--pipe" in sys.argv:  # third generation (pipeline passes) against the per-pass defaults, and factorisations that favour it
    cases = [(20, 8), (24, 1), (22, 2), (18, 16), (16, 64), (26, 1)]
    settings = [dict(big_pipe=0), dict(), dict(big_pipe=1), dict(big_pipe=1, big_factors="10,7,7"), dict(big_pipe=1, big_factors="9,9,6"),
                dict(big_pipe=1, big_factors="10,8,6"), dict(big_pipe=1, big_factors="10,6,6"), dict(big_pipe=1, big_factors="10,8"),
                dict(big_pipe=1, big_factors="9,9"), dict(big_pipe=1, big_factors="10,6"), dict(big_pipe=1, big_factors="10,9,7"),
                dict(big_pipe=1, big_factors="9,9,8"), dict(big_pipe=1, big_factors="10,10,6")]
if "--pipemask" in sys.argv:  # which passes of a three-pass transform gain from the pipeline form (big_pipe = 2 + mask)
    cases = [(24, 1), (26, 1), (22, 2)]
    settings = [dict(big_pipe=0)] + [dict(big_pipe=2 + m) for m in range(1, 8)] + [dict(big_pipe=0)]
if "--fused" in sys.argv:  # middle + last pass in one persistent launch (tile-level hand-over through the L2) against pass by pass
    cases = [(24, 1), (26, 1), (25, 1), (27, 1)]
    settings = [dict(big_fused=0), dict(big_fused=1), dict(big_fused=2), dict(big_fused=3), dict(big_fused=0)]
if "--only20" in sys.argv:
    cases = [(20, 8)]
    settings = [dict(big_v2=0), dict(big_v2=0, big_chunk=1), dict(big_v2=0, big_chunk=2), dict(big_v2=0, big_chunk=4)]
for log2n, frames in cases:
    n = 1 << log2n
    plan = ctx.plan(n, F64)
    g = torch.Generator(device=dev).manual_seed(1)
    re = torch.rand((frames, n), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    im = torch.rand((frames, n), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    ore, oim = torch.empty_like(re), torch.empty_like(re)
    ref = torch.fft.fft(torch.complex(re[0], im[0]))  # cross-check only (torch is not on the product path)
    for s in settings:
        if frames == 1 and s.get("big_chunk", 0) > 1:
            continue
        if s.get("big_factors") and sum(int(x) for x in s["big_factors"].split(",")) != log2n:
            continue
        # the pass structure is fixed when a plan is first used: forced factors get a context (and plan cache) of their own
        cx = _lib.Context(0) if s.get("big_factors") else ctx
        for k in KEYS:
            cx.tune(k, s.get(k))
        plan = cx.plan(n, F64)

        def go():
            check(L.pdsp_fft_complex_dev(plan, vp(re), vp(im), frames, vp(ore), vp(oim), 0, C.c_void_p(st.cuda_stream)))
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        err = float((torch.complex(ore[0], oim[0]) - ref).norm() / ref.norm())
        t0 = time.perf_counter()
        for _ in range(5):
            go()
        torch.cuda.synchronize()
        est = (time.perf_counter() - t0) / 5
        reps = max(10, int(0.1 / est))
        for _ in range(reps):
            go()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            go()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        frac = 32.0 * n * frames / (ms * 1e-3) / 1e9 / peak
        print(f"2^{log2n} x{frames} {s}: {ms:.4f} ms/step  {ms / frames * 1e3:.2f} us/transform  frac {frac:.3f}  rel-L2 {err:.2e}", flush=True)
        if cx is not ctx:
            torch.cuda.synchronize()
            cx.close()
    for k in KEYS:
        ctx.tune(k, None)
