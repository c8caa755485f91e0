#!/usr/bin/env python3
"""A/B of context tunables on the GPU: for each workload shape and each setting of one tunable (pdsp_ctx_tune), time
back-to-back launches after a warm-up long enough to reach the sustained clocks, and compare the outputs bit for bit.

    python scripts/ab_tune.py --key staged --values 0,1 --workloads north_star,c2,c5 [--frames 262144] [--ms 150]
    python scripts/ab_tune.py --lib pragma_dsp_b200/libexp.so ...        # time another build of the library

Prints one JSON line per (workload, value): ms per launch, frames/s, fraction of the measured HBM peak.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--key", default="staged")
    ap.add_argument("--values", default="0,1")
    ap.add_argument("--workloads", default="north_star,c2,c5,spectrum_f64")
    ap.add_argument("--frames", type=int, default=1 << 18)
    ap.add_argument("--ms", type=float, default=150.0, help="timed region per setting (after a warm-up of the same length)")
    ap.add_argument("--lib", default=None)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    from pragma_dsp_b200 import _lib
    if a.lib:
        _lib.LIB_PATH = os.path.abspath(a.lib)
    from pragma_dsp_b200._lib import F32, F64, SIDES, WINDOWS, SpectrumDesc, check, lib
    ctx = _lib.Context(0)
    L = lib()
    peak, _ = bench.measured_hbm_peak()
    dev = torch.device("cuda", 0)
    for wl in a.workloads.split(","):
        w = dict(bench.WORKLOADS[wl])
        n, frames = w["n"], a.frames
        hop = w.get("hop", n)
        if hop != n:
            frames = min(frames, w["frames"])
        prec = F64 if w["prec"] == "f64" else F32
        tdt = torch.float64 if prec == F64 else torch.float32
        sdt = torch.float64 if w["sdtype"] == "f64" else torch.float32
        x = bench.synth_frames_torch(torch, frames, n, sdt, dev, 1337) if hop == n else \
            bench.synth_stream_torch(torch, (frames - 1) * hop + n, sdt, dev, 1337)
        bins = n // 2 + 1
        amp = torch.empty((frames, bins), dtype=tdt, device=dev) if "amplitude" in w["outputs"] else None
        ph = torch.empty((frames, bins), dtype=tdt, device=dev) if "phase" in w["outputs"] else None
        pk = torch.zeros((frames, 32), dtype=torch.uint8, device=dev) if "peak" in w["outputs"] else None
        plan = ctx.plan(n, prec)
        d = SpectrumDesc(sample_dtype=F64 if w["sdtype"] == "f64" else F32, frame_len=n, hop=hop, batch=frames,
                         window=WINDOWS[w["window"]], sides=SIDES["one"], sample_rate=48000.0, raw_magnitude=0)
        st = torch.cuda.Stream(device=dev)
        bpf = bench.algorithmic_bytes_per_frame(w)
        vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        ref = None
        for v in a.values.split(","):
            ctx.tune(a.key, v if v != "default" else None)

            def go():
                check(L.pdsp_spectrum_dev(plan, C.byref(d), vp(x), vp(amp), vp(ph), vp(pk), C.c_void_p(st.cuda_stream)))
            # estimate the launch time, then warm up and time for ~a.ms each
            for _ in range(3):
                go()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                go()
            torch.cuda.synchronize()
            est = (time.perf_counter() - t0) / 5
            reps = max(5, int(a.ms * 1e-3 / est))
            for _ in range(reps):
                go()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                go()
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            fps = frames / (ms * 1e-3)
            outs = [t.clone() for t in (amp, ph, pk) if t is not None]
            same = None
            if ref is None:
                ref = outs
            else:
                same = all(bool(torch.equal(p_, q_)) for p_, q_ in zip(ref, outs))
            print(json.dumps({"tag": a.tag, "workload": wl, a.key: v, "frames": frames, "reps": reps, "ms": ms, "frames_per_s": fps,
                              "frac_of_measured_hbm": fps * bpf / 1e9 / peak, "bit_identical_to_first": same}), flush=True)
        ctx.tune(a.key, None)


if __name__ == "__main__":
    main()
