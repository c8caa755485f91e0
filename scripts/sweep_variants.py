#!/usr/bin/env python3
"""Rank the tuning variants of the N=1024 kernels (fft_config.h PDSP_VARIANT) on the GPU.

    python scripts/sweep_variants.py [--frames 65536] [--reps 30] > gpurun_out/sweep.json

For each workload shape (bench.py WORKLOADS) and variant: `reps` back-to-back launches timed with
CUDA events after 5 warm-ups; prints frames/s, achieved algorithmic GB/s and fraction of the
measured HBM peak.  Inputs+outputs exceed L2 at the default size.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from pragma_dsp_b200 import _lib  # noqa: E402
from pragma_dsp_b200._lib import F32, F64, SIDES, WINDOWS, SpectrumDesc, check, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,7,8,9,10,11,12,13")
    ap.add_argument("--workloads", default="c2,north_star,c5")
    a = ap.parse_args()
    ctx = _lib.Context(0)
    L = lib()
    peak, _ = bench.measured_hbm_peak()
    dev = torch.device("cuda", 0)
    rows = []
    for wl in a.workloads.split(","):
        w = dict(bench.WORKLOADS[wl])
        n, frames = w["n"], a.frames
        prec = F64 if w["prec"] == "f64" else F32
        tdt = torch.float64 if prec == F64 else torch.float32
        sdt = torch.float64 if w["sdtype"] == "f64" else torch.float32
        x = bench.synth_frames_torch(torch, frames, n, sdt, dev, 1337)
        bins = n // 2 + 1
        amp = torch.empty((frames, bins), dtype=tdt, device=dev) if "amplitude" in w["outputs"] else None
        pk = torch.zeros((frames, 32), dtype=torch.uint8, device=dev) if "peak" in w["outputs"] else None
        plan = ctx.plan(n, prec)
        d = SpectrumDesc(sample_dtype=F64 if w["sdtype"] == "f64" else F32, frame_len=n, hop=n, batch=frames,
                         window=WINDOWS[w["window"]], sides=SIDES["one"], sample_rate=48000.0, raw_magnitude=0)
        st = torch.cuda.Stream(device=dev)  # explicit: handle 0 would mean 'the context's own stream' to the C-ABI
        bpf = bench.algorithmic_bytes_per_frame(w)
        ref_amp = ref_pk = None
        for v in [int(s) for s in a.variants.split(",")]:
            ctx.tune("variant", v)

            def go():
                check(L.pdsp_spectrum_dev(plan, C.byref(d), C.c_void_p(x.data_ptr()),
                                          C.c_void_p(amp.data_ptr()) if amp is not None else None, None,
                                          C.c_void_p(pk.data_ptr()) if pk is not None else None, C.c_void_p(st.cuda_stream)))
            for _ in range(5):
                go()
            torch.cuda.synchronize()
            assert st.cuda_stream != 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(a.reps):
                go()
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            fps = frames / (ms * 1e-3)
            # variants must agree (rounding aside)
            agree = None
            if v == 0:
                ref_amp = amp.clone() if amp is not None else None
                ref_pk = pk.clone() if pk is not None else None
            else:
                agree = {}
                if amp is not None:
                    agree["amp_max_abs_diff"] = float((amp - ref_amp).abs().max())
                if pk is not None:
                    agree["peak_index_equal"] = bool((pk[:, :4] == ref_pk[:, :4]).all())
            row = {"workload": wl, "variant": v, "ms": ms, "frames_per_s": fps, "gbs": fps * bpf / 1e9,
                   "frac_of_measured_hbm": fps * bpf / 1e9 / peak, "agree_with_v0": agree}
            rows.append(row)
            print(json.dumps(row), flush=True)
    ctx.tune("variant", None)


if __name__ == "__main__":
    main()
