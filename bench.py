#!/usr/bin/env python3
"""bench.py - headline benchmark of the pragma-dsp FFT/spectrum hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Metric (BASELINE.json): batched N=1024 FFT frames/s (+ achieved HBM GB/s).  A "step" is one launch of the fused
kernel (frame build + window + FFT + magnitude / amplitude / peak) over one batch of synthetic multi-tone frames.
Default workload = BASELINE.json's stated target ("north_star"): fp64 Hann-windowed FFT + one-sided magnitude,
N=1024, 2^20 frames per GPU and step - 16 x the 65,536-frame batch of configs[1], so that the K timed steps
cover >= 50 ms at the clocks the part sustains under its power cap (a 65,536-frame launch lasts 0.17 ms; its
back-to-back "burst" rate is reported as a secondary key).  Weak scaling: every rank owns a batch of that size;
with N>1 and a peak output every rank also receives all ranks' per-frame peaks - by default through NVLink peer
stores fused into the kernel's epilogue (--gather p2p), alternatively a side-stream NCCL all_gather.
Other workloads: "c2" (configs[1]: spectrum() fp32 amplitude + peak), "c5" (configs[4]: 2^20 frames, fp64
window + FFT + peak only, frame-sharded; --scaling strong splits the 2^20 frames over the ranks), "c3", "c4_*",
"run_ts_*", "c1" (single-call latency of Radix2Fft.forward / spectrum()).

One JSON line on stdout (rank 0).  `value` = whole-job frames/s with inputs resident in HBM;
`e2e` = the same metric through the public host API (pdsp_spectrum) from pinned host buffers, H2D and D2H inside
the timed region.  `--impl reference` times the CPU oracle port of the reference algorithm (the reference is
TypeScript and cannot run here) on all host threads over a bounded sample of the same workload, computing the
same outputs as the GPU arm.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 1337
_STDOUT_GUARD = None


def emit_line(obj):
    """The ONE JSON line, on the real stdout."""
    global _STDOUT_GUARD
    if _STDOUT_GUARD is not None:
        _STDOUT_GUARD.restore()
        _STDOUT_GUARD = None
    print(json.dumps(obj), flush=True)

WORKLOADS = {
    # name: (precision, sample dtype, window, outputs, frames/GPU, N, description)
    "north_star": dict(prec="f64", sdtype="f64", window="hann", outputs=("amplitude",), frames=1 << 20, n=1024,
                       desc="BASELINE target: fp64 Hann-windowed FFT + one-sided magnitude, N=1024, 2^20 frames per step "
                            "(16 x the 65,536-frame batch of configs[1]: 20 steps = a sustained >= 50 ms region)"),
    "c2": dict(prec="f32", sdtype="f32", window="hann", outputs=("amplitude", "peak"), frames=1 << 21, n=1024,
               desc="configs[1] spectrum() batched, N=1024 fp32, Hann, one-sided amplitude + peak @48kHz; 2^21 frames per step "
                    "(32 x 65,536)"),
    # BASELINE config C3: STFT via spectrumStream - 10 min @ 48 kHz, N=4096, hop 1024, Hann, magnitude + phase,
    # Float32Array frames (spectrumStream's element type) computed in fp64 like the reference
    "c3": dict(prec="f64", sdtype="f32", window="hann", outputs=("amplitude", "phase"), frames=28122, n=4096, hop=1024,
               desc="STFT: 28,800,000 fp32 samples (10 min @48kHz), N=4096 hop 1024 Hann, fp64 amplitude + phase, 28122 frames"),
    # spectrum()'s own output set (amplitude + phase + peak, src/public/spectrum.ts:121-134), batched
    "spectrum_f64": dict(prec="f64", sdtype="f64", window="hann", outputs=("amplitude", "phase", "peak"), frames=1 << 19, n=1024,
                         desc="spectrum() default outputs (amplitude + phase + peak), fp64 N=1024 Hann, 2^19 frames per step"),
    # bench/run.ts's own workloads: FFT.forward(input, out) on real fp64 frames, all N bins out
    "run_ts_2048": dict(prec="f64", sdtype="f64", window="rect", outputs=("complex",), frames=32768, n=2048, kind="r2c_forward",
                        desc="bench/run.ts shape: FFT.forward(input, out), N=2048 real fp64 -> N complex bins, 32768 frames"),
    "run_ts_4096": dict(prec="f64", sdtype="f64", window="rect", outputs=("complex",), frames=16384, n=4096, kind="r2c_forward",
                        desc="bench/run.ts shape: FFT.forward(input, out), N=4096 real fp64 -> N complex bins, 16384 frames"),
    # BASELINE config C4: large single complex fp64 transforms (multi-pass path); a "frame" is one transform
    "c4_2e20": dict(prec="f64", sdtype="f64", window="rect", outputs=("complex",), frames=8, n=1 << 20, kind="c2c",
                    desc="complex fp64 FFT, N=2^20, 8 transforms per step (multi-pass 1024x1024)"),
    "c4_2e24": dict(prec="f64", sdtype="f64", window="rect", outputs=("complex",), frames=1, n=1 << 24, kind="c2c",
                    desc="complex fp64 FFT, N=2^24 (multi-pass 256x256x256)"),
    # BASELINE config C5 as written: 2^20 frames x N=1024, window + FFT + peak argmax, peaks gathered on every rank
    "c5": dict(prec="f64", sdtype="f64", window="hann", outputs=("peak",), frames=1 << 20, n=1024,
               desc="configs[4]: 2^20 frames x N=1024 fp64 window + FFT + peak argmax only, frame-sharded, peaks gathered"),
    "c1": dict(prec="f64", sdtype="f64", window="rect", outputs=("complex",), frames=1, n=1024, kind="latency",
               desc="configs[0]: Radix2Fft.forward N=1024 fp64 on one frame of sin(i) (README core example) - us per call"),
}
BURST_FRAMES = 65536  # the batch of BASELINE configs[1]: secondary "burst" timing of the same kernel


def algorithmic_bytes_per_frame(w) -> int:
    """compulsory HBM traffic per frame; cached tables excluded."""
    es = 8 if w["sdtype"] == "f64" else 4
    os_ = 8 if w["prec"] == "f64" else 4
    if w.get("kind") == "c2c":
        return 2 * 2 * es * w["n"]  # both planes in, both planes out
    if w.get("kind") == "r2c_forward":
        return es * w["n"] + 2 * os_ * w["n"]  # real frame in, both planes (all N bins) out
    bins = w["n"] // 2 + 1
    b = w.get("hop", w["n"]) * es  # unique input bytes per frame (overlap re-reads are expected to hit L2)
    if "amplitude" in w["outputs"]:
        b += bins * os_
    if "phase" in w["outputs"]:
        b += bins * os_
    if "peak" in w["outputs"]:
        b += 32 if w["prec"] == "f64" else 16
    return b


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, w):
    """The reference's own CPU implementation of the path, restated in C (oracle/), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle

    threads = len(os.sched_getaffinity(0))  # torchrun pins OMP_NUM_THREADS=1; the oracle takes an explicit count
    n = w["n"]
    if w.get("kind") == "latency":
        po = oracle.FFT(n)
        x = np.sin(np.arange(n, dtype=np.float64))
        calls = max(200, args.steps * 50)
        for _ in range(20):
            po.forward(x)
        t0 = time.perf_counter()
        for _ in range(calls):
            po.forward(x)
        us = (time.perf_counter() - t0) / calls * 1e6
        emit_line({"impl": "reference", "metric": "us_per_call", "value": us, "unit": "us", "n_gpus": args.gpus, "steps": calls,
                   "warmup": 20, "ms_per_step": us / 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
                   "dtype": "f64", "data": "synthetic", "config": {"workload": args.workload, "description": w["desc"], "fft_size": n},
                   "cpu_baseline": {"value": us, "unit": "us", "cores": 1, "kind": "port",
                                    "sample": "Radix2Fft.forward on one frame of sin(i) through oracle/pragma_oracle.c, one core"},
                   "e2e": {"value": us, "unit": "us", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return 0
    if w.get("kind") in ("c2c", "r2c_forward"):
        rng = np.random.default_rng(SEED)
        nf = 1 if w["kind"] == "c2c" else min(w["frames"], 2048 * threads)
        re, im = rng.uniform(-1, 1, (nf, n)), rng.uniform(-1, 1, (nf, n))
        plan = oracle.FFT(n)
        run = (lambda: plan.forwardComplex(re, im)) if w["kind"] == "c2c" else (lambda: plan.forward(re, threads=threads))
        for _ in range(min(args.warmup, 1)):
            run()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run()
        dt = time.perf_counter() - t0
        fps = nf * args.steps / dt
        emit_line({"impl": "reference", "metric": "frames_per_s", "value": fps, "unit": "frames/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": args.workload, "description": w["desc"], "fft_size": n},
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1 if w["kind"] == "c2c" else threads, "kind": "port",
                                           "sample": "one transform per step (a single radix-2 transform does not thread)"
                                           if w["kind"] == "c2c" else f"{nf} frames per step, frames split over threads"},
                          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return 0
    # bounded sample: about 0.5 s of all-core work per step
    sample = int(min(w["frames"], max(2048, 12000 * threads)))
    x = synth_frames_numpy(sample, n, np.float64 if w["sdtype"] == "f64" else np.float32)
    outs = tuple(w["outputs"])
    hop = w.get("hop", n)
    if hop != n:  # STFT: overlapping views of one stream
        x = np.ascontiguousarray(x.reshape(-1)[:(sample - 1) * hop + n])
        kw = dict(frameLen=n, hop=hop, batch=sample)
    else:
        kw = {}

    def step():
        # the outputs the GPU arm produces, nothing more: lean=True runs Math.atan2 only where an output needs it
        oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window=w["window"], sides="one",
                              want_amplitude="amplitude" in outs, want_phase="phase" in outs, want_peaks="peak" in outs,
                              threads=threads, lean=True, **kw)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    what = "window, radix-2 FFT, hypot, scaling" + (", atan2 per bin" if "phase" in outs else "") + \
           (", findPeak + atan2 at the peak bin" if "peak" in outs and "phase" not in outs else "") + \
           (", findPeak" if "peak" in outs and "phase" in outs else "")
    line = {
        "impl": "reference", "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, w),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} frames/step of the same synthetic workload, same outputs as the GPU arm "
                                   f"({', '.join(outs)}): oracle/pragma_oracle.c (op-for-op C port of src/core/fft.ts + "
                                   f"spectrum.ts: {what}), gcc -O2 -ffp-contract=off, OpenMP static split"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)
    return 0


def workload_config(args, w):
    return {"workload": args.workload, "description": w["desc"], "fft_size": w["n"], "hop": w.get("hop", w["n"]),
            "frames_per_gpu": w["frames"],
            "window": w["window"], "sides": "one", "sample_rate": 48000, "outputs": list(w["outputs"]),
            "sample_dtype": w["sdtype"], "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)",
            "parallelism": f"frames sharded x{args.gpus}", "scaling": getattr(args, "scaling", "weak")}


# ----------------------------------------------------------------------------- synthetic data
def synth_frames_numpy(frames, n, dtype, seed=SEED):
    """Multi-tone frames per SURVEY 8(d): 3 tones, dominant A1=1, off-bin by at most a quarter bin."""
    rng = np.random.default_rng(seed)
    out = np.empty((frames, n), dtype=dtype)
    t = np.arange(n, dtype=np.float64)
    for lo in range(0, frames, 4096):
        hi = min(frames, lo + 4096)
        b = hi - lo
        k = rng.integers(8, n // 2 - 8, size=(b, 3)) + rng.uniform(-0.25, 0.25, size=(b, 3))
        a = np.concatenate([np.ones((b, 1)), rng.uniform(0.1, 0.5, size=(b, 2))], axis=1)
        ph = rng.uniform(0, 2 * np.pi, size=(b, 3))
        acc = np.zeros((b, n))
        for j in range(3):
            acc += a[:, j, None] * np.sin(2 * np.pi * k[:, j, None] * t[None, :] / n + ph[:, j, None])
        out[lo:hi] = acc.astype(dtype)
    return out


def synth_stream_torch(torch, samples, dtype, device, seed):
    """One long multi-tone stream (3 tones, slowly chirped) for the STFT workload."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(samples, dtype=dtype, device=device)
    f0 = torch.tensor([440.0, 1500.0, 5200.0], dtype=torch.float64, device=device)
    a = torch.tensor([1.0, 0.4, 0.2], dtype=torch.float64, device=device)
    for lo in range(0, samples, 1 << 22):
        hi = min(samples, lo + (1 << 22))
        t = torch.arange(lo, hi, dtype=torch.float64, device=device) / 48000.0
        acc = torch.zeros(hi - lo, dtype=torch.float64, device=device)
        for j in range(3):
            acc += a[j] * torch.sin(2 * np.pi * (f0[j] * t + 0.5 * (20.0 * (j + 1)) * t * t / 60.0))
        out[lo:hi] = acc.to(dtype)
    return out


def synth_frames_torch(torch, frames, n, dtype, device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((frames, n), dtype=dtype, device=device)
    t = torch.arange(n, dtype=torch.float64, device=device)
    for lo in range(0, frames, 8192):
        hi = min(frames, lo + 8192)
        b = hi - lo
        k = torch.randint(8, n // 2 - 8, (b, 3), generator=g, device=device).double()
        k += torch.rand((b, 3), generator=g, device=device, dtype=torch.float64) * 0.5 - 0.25
        a = torch.cat([torch.ones((b, 1), device=device, dtype=torch.float64),
                       torch.rand((b, 2), generator=g, device=device, dtype=torch.float64) * 0.4 + 0.1], dim=1)
        ph = torch.rand((b, 3), generator=g, device=device, dtype=torch.float64) * (2 * np.pi)
        acc = torch.zeros((b, n), dtype=torch.float64, device=device)
        for j in range(3):
            acc += a[:, j, None] * torch.sin(2 * np.pi * k[:, j, None] * t[None, :] / n + ph[:, j, None])
        out[lo:hi] = acc.to(dtype)
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = get_reasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- b200 arm
def _traffic_of(workload, frames):
    """DRAM bytes per launch of the workload's dominant kernel: the per-frame figure of the committed `ncu --set full`
    capture (profiles/traffic.json names the capture) scaled to this launch's frame count.  (None, None) if not captured."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        e = json.load(open(tp)).get(workload)
        if isinstance(e, dict):
            return e["dram_bytes_per_frame"] * frames, e.get("source")
    except Exception:
        pass
    return None, None


def run_b200(args, w):
    import torch
    import torch.distributed as dist

    from pragma_dsp_b200 import _lib, spectrum_batch
    from pragma_dsp_b200._lib import F32, F64, PEAK_F32, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side barrier: waiting ranks must not keep a kernel spinning on their GPU
    ctx = _lib.Context(local)
    n, frames = w["n"], args.frames or w["frames"]
    if args.scaling == "strong":  # total work fixed: the workload's frames are split over the ranks
        frames = -(-frames // world)
    prec = F64 if w["prec"] == "f64" else F32
    tdt = torch.float64 if w["prec"] == "f64" else torch.float32
    sdt = torch.float64 if w["sdtype"] == "f64" else torch.float32
    plan = ctx.plan(n, prec)
    bins = n // 2 + 1
    pk_bytes = 32 if prec == F64 else 16

    hop = w.get("hop", n)
    if hop == n:
        x = synth_frames_torch(torch, frames, n, sdt, dev, SEED + rank)
    else:  # one long stream, frames are overlapping views
        x = synth_stream_torch(torch, (frames - 1) * hop + n, sdt, dev, SEED + rank)
    amp = torch.empty((frames, bins), dtype=tdt, device=dev) if "amplitude" in w["outputs"] else None
    ph = torch.empty((frames, bins), dtype=tdt, device=dev) if "phase" in w["outputs"] else None
    want_peak = "peak" in w["outputs"]
    peaks = [torch.zeros((frames, pk_bytes), dtype=torch.uint8, device=dev) for _ in range(2)] if want_peak else None
    gather = args.gather if (want_peak and world > 1) else "none"
    gathered = [torch.zeros((world * frames, pk_bytes), dtype=torch.uint8, device=dev) for _ in range(2)] \
        if gather == "nccl" else None
    peers = gbuf = None
    if gather == "p2p":
        # every rank owns a (world * frames) record buffer; all ranks map all of them (CUDA IPC over NVLink)
        # and the kernel's epilogue stores each finished record into every one of them
        gbuf = C.c_void_p()
        check(lib().pdsp_dev_alloc(ctx.h, world * frames * pk_bytes, C.byref(gbuf)))
        hbuf = C.create_string_buffer(64)
        check(lib().pdsp_ipc_export(ctx.h, gbuf, hbuf))
        handles = [None] * world
        dist.all_gather_object(handles, hbuf.raw)
        peers = (C.c_void_p * 8)()
        for g_ in range(world):
            if g_ == rank:
                peers[g_] = gbuf.value
            else:
                pp = C.c_void_p()
                check(lib().pdsp_ipc_open(ctx.h, handles[g_], C.byref(pp)))
                peers[g_] = pp.value
        dist.barrier()
    desc = SpectrumDesc(sample_dtype=F64 if w["sdtype"] == "f64" else F32, frame_len=n, hop=hop, batch=frames,
                        window=WINDOWS[w["window"]], sides=SIDES["one"], sample_rate=48000.0, raw_magnitude=0)
    compute = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev) if gathered else None
    torch.cuda.synchronize(dev)  # inputs and zero fills were produced on torch's stream: done before anything runs on `compute`
    L = lib()

    def vp(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    kernel_events = []

    def step(i, timed=False):
        pb = peaks[i & 1] if want_peak else None
        with torch.cuda.stream(compute):
            if comm is not None and i >= 2:
                compute.wait_event(gather_done[i & 1])  # the gather that last read this peaks buffer
            # per-launch duration (roofline.achieved): CUDA events around every launch of the timed region (a launch
            # lasts milliseconds at these batch sizes: the event pair's few microseconds do not show)
            if timed:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(compute)
            if peers is not None:
                check(L.pdsp_spectrum_dev_gather(plan, C.byref(desc), vp(x), vp(amp), vp(ph), vp(pb), peers, world,
                                                 rank * frames, C.c_void_p(compute.cuda_stream)))
            else:
                check(L.pdsp_spectrum_dev(plan, C.byref(desc), vp(x), vp(amp), vp(ph), vp(pb), C.c_void_p(compute.cuda_stream)))
            if timed:
                e1.record(compute)
                kernel_events.append((e0, e1))
            if comm is not None:
                ev = torch.cuda.Event()
                ev.record(compute)
                with torch.cuda.stream(comm):
                    comm.wait_event(ev)
                    dist.all_gather_into_tensor(gathered[i & 1], pb)
                    gather_done[i & 1].record(comm)

    gather_done = [torch.cuda.Event(), torch.cuda.Event()]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- secondary "burst" figure: 20 back-to-back launches of the 65,536-frame batch from a cool start (~3 ms at
    # boost clocks) - what round 1 reported as `value`
    burst = None
    if frames > BURST_FRAMES and not args.quick and hop == n:
        bdesc = SpectrumDesc(sample_dtype=desc.sample_dtype, frame_len=n, hop=hop, batch=BURST_FRAMES, window=desc.window,
                             sides=desc.sides, sample_rate=48000.0, raw_magnitude=0)

        def burst_launch():
            check(L.pdsp_spectrum_dev(plan, C.byref(bdesc), vp(x), vp(amp), vp(ph), vp(peaks[0]) if want_peak else None,
                                      C.c_void_p(compute.cuda_stream)))
        for _ in range(3):
            burst_launch()
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(compute)
        for _ in range(20):
            burst_launch()
        b1.record(compute)
        barrier()
        burst = {"frames_per_launch": BURST_FRAMES, "launches": 20, "ms_per_launch": b0.elapsed_time(b1) / 20}

    # ---- warm-up: at least W steps AND at least 0.2 s, so that the timed region runs at the clocks the part sustains
    # under its power cap rather than at the first milliseconds' boost; then EXACTLY K timed steps, device time,
    # max over ranks
    t0 = time.perf_counter()
    i = 0
    while i < max(3, args.warmup) or (not args.quick and time.perf_counter() - t0 < 0.2):
        step(i)
        i += 1
        if i % 8 == 0:
            compute.synchronize()
    warm_steps = i
    barrier()
    launches0 = ctx.launch_count
    if sampler:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(compute)
    for i in range(args.steps):
        step(i, timed=True)
    if comm is not None:
        compute.wait_stream(comm)
    t_end.record(compute)
    barrier()
    if sampler:
        sampler.stop()
    launches = ctx.launch_count - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    if world > 1:
        tt = torch.tensor([elapsed_ms, kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms, kern_ms = float(tt[0]), float(tt[1])
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt[0])
    barrier()

    gather_check = None
    if peers is not None:
        # EVERY segment of this rank's gathered buffer must be byte-equal to what the owning rank computed: all-gather
        # the ranks' local record buffers with NCCL (outside every timed region) and compare the whole buffer
        mine = peaks[(args.steps - 1) & 1]
        ref = torch.empty((world * frames, pk_bytes), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(ref, mine)
        host = torch.empty((world * frames, pk_bytes), dtype=torch.uint8).pin_memory()
        check(L.pdsp_memcpy_d2h(ctx.h, C.c_void_p(host.data_ptr()), gbuf, host.numel(), C.c_void_p(compute.cuda_stream)))
        compute.synchronize()
        ref_h = ref.cpu()
        seg_ok = [bool(torch.equal(host[g_ * frames:(g_ + 1) * frames], ref_h[g_ * frames:(g_ + 1) * frames])) for g_ in range(world)]
        ok = torch.tensor([int(all(seg_ok))], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        gather_check = {"mode": "p2p", "segments_byte_equal_on_rank0": seg_ok, "all_ranks_all_segments_equal": bool(int(ok[0])),
                        "checked_against": "NCCL all_gather of every rank's local records"}
        del ref, ref_h, host
        barrier()
        for g_ in range(world):
            if g_ != rank:
                check(L.pdsp_ipc_close(ctx.h, C.c_void_p(peers[g_])))
        barrier()
        check(L.pdsp_dev_free(ctx.h, gbuf))
    elif gathered is not None:
        mine = peaks[(args.steps - 1) & 1]
        got = gathered[(args.steps - 1) & 1]
        gather_check = {"mode": "nccl", "own_segment_equal": bool(torch.equal(got[rank * frames:(rank + 1) * frames], mine))}

    # ---- e2e: public host API, pinned host buffers, H2D + D2H inside the timed region
    e2e_frames = min(frames, 131072) if hop == n else frames
    e2e_steps = max(2, min(args.steps, 10)) if not args.quick else 1
    xs = x[:e2e_frames] if hop == n else x
    hx = torch.empty(tuple(xs.shape), dtype=sdt).pin_memory()
    hx.copy_(xs.cpu())
    hx_np = hx.numpy()
    outs = tuple(w["outputs"])
    h2d = hx.numel() * hx.element_size()
    d2h = e2e_frames * ((bins * (8 if prec == F64 else 4) if "amplitude" in outs else 0) +
                        (bins * (8 if prec == F64 else 4) if "phase" in outs else 0) + (pk_bytes if want_peak else 0))
    # pinned output buffers handed to the C-ABI directly (what createComplexArray-style pinned arrays give JS)
    h_amp = torch.empty((e2e_frames, bins), dtype=tdt).pin_memory() if "amplitude" in outs else None
    h_ph = torch.empty((e2e_frames, bins), dtype=tdt).pin_memory() if "phase" in outs else None
    h_pk = torch.empty((e2e_frames, pk_bytes), dtype=torch.uint8).pin_memory() if want_peak else None
    edesc = SpectrumDesc(sample_dtype=desc.sample_dtype, frame_len=n, hop=hop, batch=e2e_frames, window=desc.window,
                         sides=desc.sides, sample_rate=48000.0, raw_magnitude=0)

    def e2e_step():
        check(L.pdsp_spectrum(plan, C.byref(edesc), C.c_void_p(hx.data_ptr()),
                              C.c_void_p(h_amp.data_ptr()) if h_amp is not None else None,
                              C.c_void_p(h_ph.data_ptr()) if h_ph is not None else None,
                              C.c_void_p(h_pk.data_ptr()) if h_pk is not None else None))

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])
    e2e_fps = e2e_frames * world * e2e_steps / e2e_s
    e2e_api = "pdsp_spectrum (host pinned buffers)"
    e2e_per_rank = None
    if world > 1 and hop == n and not args.quick:
        # N GPUs through ONE host call: rank 0 drives all of the box's GPUs with the library's device group
        # (pdsp_group_spectrum: one block of frames per device, each through its device's staging pipeline on a NUMA-pinned
        # host thread) while the other ranks wait.  This is what a user of the library calls; the per-rank figure above
        # (N processes, one pdsp_spectrum each) is kept beside it.
        e2e_per_rank = {"value": e2e_fps, "api": "N processes x pdsp_spectrum, concurrently"}

        def barrier_cpu():
            # an NCCL barrier keeps a kernel polling on every waiting rank's GPU, and rank 0's group work on that GPU would
            # then be time-sliced against it (measured: half the throughput); the waiting ranks block on the host instead
            torch.cuda.synchronize(dev)
            dist.barrier(group=cpu_group)
        barrier_cpu()
        if rank == 0:
            from pragma_dsp_b200 import DeviceGroup
            gx = torch.empty((world * e2e_frames, n), dtype=sdt).pin_memory()
            for r_ in range(world):
                gx[r_ * e2e_frames:(r_ + 1) * e2e_frames].copy_(hx)
            g_amp = torch.empty((world * e2e_frames, bins), dtype=tdt).pin_memory() if "amplitude" in outs else None
            g_ph = torch.empty((world * e2e_frames, bins), dtype=tdt).pin_memory() if "phase" in outs else None
            g_pk = torch.empty((world * e2e_frames, pk_bytes), dtype=torch.uint8).pin_memory() if want_peak else None
            gdesc = SpectrumDesc(sample_dtype=desc.sample_dtype, frame_len=n, hop=hop, batch=world * e2e_frames, window=desc.window,
                                 sides=desc.sides, sample_rate=48000.0, raw_magnitude=0)
            grp = DeviceGroup(list(range(world)))
            dp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731

            def group_step():
                check(L.pdsp_group_spectrum(grp._h, n, prec, C.byref(gdesc), dp(gx), dp(g_amp), dp(g_ph), dp(g_pk)))
            for _ in range(2):
                group_step()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                group_step()
            g_s = time.perf_counter() - t0
            grp.close()
            g_fps = world * e2e_frames * e2e_steps / g_s
            if h_amp is not None:  # the blocks were copies of rank 0's batch: same results expected in every block
                assert torch.equal(g_amp[:e2e_frames], g_amp[(world - 1) * e2e_frames:]), "group blocks disagree"
            e2e_fps, e2e_api = g_fps, "pdsp_group_spectrum (one host call, all GPUs, host pinned buffers)"
            del gx, g_amp, g_ph, g_pk
        barrier_cpu()

    # ---- ingestion ring: 65,536 of the frames pushed in blocks of 256 from pageable memory, results
    # popped in order - what a frame-at-a-time source (spectrumStream) sees
    ingest = None
    if rank == 0 and world == 1 and not args.quick and hop == n:
        from pragma_dsp_b200 import IngestRing
        ing_frames = min(e2e_frames, 65536)
        src = np.array(hx_np[:ing_frames], copy=True)  # pageable, like a JS typed array
        ring = IngestRing(n, sampleRate=48000.0, fftSize=n, window=w["window"], precision=w["prec"],
                          sample_dtype=src.dtype, outputs=outs, framesPerChunk=4096, depth=3, context=ctx)

        pdt = PEAK_F64 if prec == F64 else PEAK_F32
        odt = np.float64 if prec == F64 else np.float32
        o_amp = np.empty((4096, bins), dtype=odt) if "amplitude" in outs else None
        o_ph = np.empty((4096, bins), dtype=odt) if "phase" in outs else None
        o_pk = np.zeros(4096, dtype=pdt) if want_peak else None

        def ingest_pass(source, push):
            pushed = popped = 0
            while pushed < ing_frames:
                k = push(source[pushed:pushed + 256])
                pushed += k
                if k == 0:
                    popped += ring.pop_into(o_amp, o_ph, o_pk)
            ring.flush()
            while popped < ing_frames:
                popped += ring.pop_into(o_amp, o_ph, o_pk)

        ingest_pass(src, ring.push)
        t0 = time.perf_counter()
        ingest_pass(src, ring.push)
        ingest_s = time.perf_counter() - t0
        # the same frames from pinned memory (what hostAlloc-backed typed arrays give a JS caller): no host copy on the way in
        pin_src = hx_np[:ing_frames]
        ingest_pass(pin_src, ring.push_pinned)
        t0 = time.perf_counter()
        ingest_pass(pin_src, ring.push_pinned)
        ingest_pin_s = time.perf_counter() - t0
        ring.close()
        ingest = {"value": ing_frames / ingest_s, "unit": "frames/s", "api": "pdsp_ingest_push/pop, 256-frame pushes, pageable source, "
                  "results popped into preallocated arrays", "pinned_source_value": ing_frames / ingest_pin_s,
                  "pinned_api": "pdsp_ingest_push_pinned (DMA straight from the caller's pinned frames)",
                  "frames_per_chunk": 4096, "depth": 3, "frames": ing_frames}

    # ---- parity spot check (outside every timed region): first 256 frames vs the oracle
    parity = None
    cpu_baseline = None
    if rank == 0:
        import oracle
        if hop == n:
            sub, kw = hx_np[:256], {}
        else:
            sub, kw = hx_np[:255 * hop + n], dict(frameLen=n, hop=hop, batch=256)
        got = spectrum_batch(sub, sampleRate=48000.0, fftSize=n, window=w["window"], precision=w["prec"], context=ctx, **kw)
        ref = oracle.spectrum_batch(sub, fftSize=n, sampleRate=48000.0, window=w["window"], **kw)
        parity = {"frames": 256, "peak_index_equal": bool((got["peaks"]["index"] == ref["peaks"]["index"]).all()),
                  "amp_max_abs_err": float(np.abs(got["amplitude"] - ref["amplitude"]).max())}
        if world == 1 and not args.quick:
            sample = min(e2e_frames, 65536 if n <= 1024 else 16384)
            if hop == n:
                sx, skw = hx_np[:sample], {}
            else:
                sx, skw = hx_np[:(sample - 1) * hop + n], dict(frameLen=n, hop=hop, batch=sample)
            okw = dict(fftSize=n, sampleRate=48000.0, window=w["window"], want_amplitude="amplitude" in outs,
                       want_phase="phase" in outs, want_peaks=want_peak, lean=True)
            t0 = time.perf_counter()
            oracle.spectrum_batch(sx, threads=1, **okw, **skw)
            one = sample / (time.perf_counter() - t0)
            thr = oracle.max_threads()
            t0 = time.perf_counter()
            oracle.spectrum_batch(sx, threads=thr, **okw, **skw)
            allc = sample / (time.perf_counter() - t0)
            cpu_baseline = {"value": one, "unit": "frames/s", "cores": 1, "kind": "port",
                            "sample": f"first {sample} frames of the same batch, same outputs ({', '.join(outs)}), through "
                                      f"oracle/pragma_oracle.c (single thread, like the single-threaded JS reference); all {thr} "
                                      f"host threads: {allc:.0f} frames/s"}

    # ---- library comparator, CONTEXT ONLY (SURVEY 8d "optional sanity comparator"): the same batch through cuFFT as a user
    # of a stock library would write it - window multiply, torch.fft.rfft (cuFFT D2Z / R2C), abs() - three library kernels and
    # three passes over HBM.  Not a product path, not a target, never part of `value`; it says what fusing is worth.
    # Opt-in (--comparator): the default run, which the driver measures, never touches a library FFT.
    comparator = None
    if args.comparator and rank == 0 and world == 1 and hop == n and tuple(w["outputs"]) == ("amplitude",):
        try:
            cb = min(frames, 1 << 18)
            wt = torch.from_numpy(__import__("oracle").createWindow(w["window"], n)).to(dev).to(tdt)
            xs = x[:cb].to(tdt)

            def lib_step():
                return torch.fft.rfft(xs * wt, dim=1).abs() * (2.0 / n)
            for _ in range(3):
                lib_step()
            torch.cuda.synchronize(dev)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            c0.record()
            for _ in range(reps):
                lib_step()
            c1.record()
            torch.cuda.synchronize(dev)
            lib_fps = cb * reps / (c0.elapsed_time(c1) * 1e-3)
            comparator = {"what": "torch.fft.rfft (cuFFT) pipeline: x * window -> rfft -> abs * 2/N, device-resident, "
                                  f"{cb} frames per call; context only, not a product path", "value": lib_fps, "unit": "frames/s",
                          "this_over_library": None}
            del xs
        except Exception as e:  # the comparator must never cost the bench line
            comparator = {"what": "torch.fft.rfft (cuFFT) pipeline", "error": str(e)[:200]}

    if rank == 0:
        total_frames = frames * world
        fps = total_frames * args.steps / (elapsed_ms * 1e-3)
        if comparator and comparator.get("value"):
            comparator["this_over_library"] = fps / comparator["value"]
        bpf = algorithmic_bytes_per_frame(w)
        peak, peak_src = measured_hbm_peak()
        achieved = bpf * frames / (kern_ms * 1e-3) / 1e9
        traffic, traffic_src = _traffic_of(args.workload, frames)
        if burst:
            burst["value"] = BURST_FRAMES / (burst["ms_per_launch"] * 1e-3)
            burst["frac"] = bpf * burst["value"] / 1e9 / peak
            burst["note"] = "boost clocks, ~3 ms: not the sustained rate"
        line = {
            "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm_steps, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": w["prec"], "data": "synthetic",
            "config": dict(workload_config(args, w), frames_per_gpu=frames),
            "hbm_gbs": fps * bpf / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "r2c_kernel",
                         "algorithmic_bytes_per_frame": bpf, "kernel_ms": kern_ms,
                         "timed": f"CUDA events around each of the {args.steps} launches of the timed region, mean"},
            "burst": burst,
            "cpu_baseline": cpu_baseline,
            "library_comparator": comparator,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d * (world if e2e_per_rank else 1),
                    "d2h_bytes_per_step": d2h * (world if e2e_per_rank else 1),
                    "steps": e2e_steps, "frames_per_step": e2e_frames * (world if e2e_per_rank else 1), "api": e2e_api,
                    "per_rank_processes": e2e_per_rank},
            "ingest": ingest,
            "gpu_launches": launches,
            "clocks": sampler.summary() if sampler else None,
            "parity": parity,
            "gather": gather_check,
        }
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


# ----------------------------------------------------------------------------- b200 arm, large transforms (C4)
def run_b200_c2c(args, w):
    """BASELINE config C4: single large complex fp64 transforms on the multi-pass path.  Replicas only
    (SURVEY 8e): with N ranks every rank transforms its own arrays; no collective."""
    import torch
    import torch.distributed as dist

    from pragma_dsp_b200 import _lib
    from pragma_dsp_b200._lib import F64, check, lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.Context(local)
    L = lib()
    n, frames = w["n"], args.frames or w["frames"]
    real_in = w.get("kind") == "r2c_forward"
    plan = ctx.plan(n, F64)
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    re = torch.rand((frames, n), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    im = None if real_in else torch.rand((frames, n), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    ore, oim = torch.empty_like(re), torch.empty_like(re)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize(dev)  # the inputs were generated on torch's stream
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def step():
        if real_in:
            check(L.pdsp_fft_forward_real_dev(plan, vp(re), F64, frames, vp(ore), vp(oim), 1, C.c_void_p(st.cuda_stream)))
        else:
            check(L.pdsp_fft_complex_dev(plan, vp(re), vp(im), frames, vp(ore), vp(oim), 0, C.c_void_p(st.cuda_stream)))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = ctx.launch_count
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    barrier()
    launches = ctx.launch_count - launches0
    elapsed_ms = e0.elapsed_time(e1)
    if sampler and not args.quick:
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.6:
            step()
            st.synchronize()
    if sampler:
        sampler.stop()
    if world > 1:
        tt = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt[0])
    # e2e: host planes (pinned) through pdsp_fft_forward_complex
    e2e_steps = 1 if args.quick else max(2, min(args.steps, 5))
    hre = re.cpu().pin_memory()
    him = None if real_in else im.cpu().pin_memory()
    hor, hoi = torch.empty_like(hre).pin_memory(), torch.empty_like(hre).pin_memory()

    def e2e_step():
        if real_in:
            check(L.pdsp_fft_forward_real(plan, vp(hre), F64, frames, vp(hor), vp(hoi)))
        else:
            check(L.pdsp_fft_forward_complex(plan, vp(hre), vp(him), frames, vp(hor), vp(hoi)))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    parity = cpu_baseline = None
    if rank == 0:
        ref = np.fft.fft(hre[0].numpy() if real_in else hre[0].numpy() + 1j * him[0].numpy())
        got = hor[0].numpy() + 1j * hoi[0].numpy()
        parity = {"frames": 1, "rel_l2_vs_numpy": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
                  "bound": 1e-12 * np.log2(n)}
        if world == 1 and not args.quick:
            import oracle
            plan_o = oracle.FFT(n)
            ns = 1 if n > 65536 else min(frames, 4096)
            t0 = time.perf_counter()
            if real_in:
                plan_o.forward(hre[:ns].numpy())
            else:
                plan_o.forwardComplex(hre[:ns].numpy(), him[:ns].numpy())
            cpu_baseline = {"value": ns / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "port",
                            "sample": f"{ns} transform(s) of the same size through oracle/pragma_oracle.c (radix-2, single thread)"}
        bpf = algorithmic_bytes_per_frame(w)
        peak, peak_src = measured_hbm_peak()
        fps = frames * world * args.steps / (elapsed_ms * 1e-3)
        step_ms = elapsed_ms / args.steps
        achieved = bpf * frames / (step_ms * 1e-3) / 1e9
        line = {
            "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "fft_size": n, "frames_per_gpu": frames,
                       "l2": "2^24: 1 GiB per step >> L2; 2^20: 8 x 32 MiB in + out per step > 126 MB L2",
                       "parallelism": f"replicas x{world}"},
            "hbm_gbs": fps * bpf / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic_of(args.workload, frames)[0], "traffic_source": _traffic_of(args.workload, frames)[1],
                         "peak_source": peak_src,
                         "kernel": "r2c_kernel (MD_CPLX)" if real_in else "bigfft_pass(_tma)_kernel (all passes of a transform)",
                         "algorithmic_bytes_per_frame": bpf, "kernel_ms": step_ms},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": frames * world * e2e_steps / e2e_s, "unit": "frames/s",
                    "h2d_bytes_per_step": (8 if real_in else 16) * n * frames, "d2h_bytes_per_step": 16 * n * frames,
                    "steps": e2e_steps, "api": "pdsp_fft_forward_real (host pinned)" if real_in else "pdsp_fft_forward_complex (host pinned)"},
            "gpu_launches": launches * world, "clocks": sampler.summary() if sampler else None, "parity": parity,
        }
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


# ----------------------------------------------------------------------------- b200 arm, single-call latency (C1)
def run_b200_latency(args, w):
    """BASELINE configs[0]: Radix2Fft.forward(input, out) on ONE N=1024 fp64 frame of sin(i) (README.md:43-50) - the
    reference's primary call shape.  Reports microseconds per call through the C ABI (pdsp_fft_forward_real, pageable
    host arrays like JS typed arrays) and through spectrum()'s entry (pdsp_spectrum, one frame), next to the 1-core
    oracle on the same frame, and the batch size at which the host entry point overtakes the CPU port."""
    import torch  # noqa: F401  (device bring-up only)

    import oracle
    from pragma_dsp_b200 import _lib
    from pragma_dsp_b200._lib import F64, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _lib.Context(local)
    L = lib()
    n = w["n"]
    plan = ctx.plan(n, F64)
    x = np.sin(np.arange(n, dtype=np.float64))
    ore, oim = np.empty(n), np.empty(n)
    vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    calls = max(200, args.steps * 50)

    def time_calls(fn, k):
        for _ in range(max(10, args.warmup * 4)):
            fn()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        return (time.perf_counter() - t0) / k * 1e6

    ping_us = time_calls(lambda: check(L.pdsp_ctx_ping(ctx.h)), calls)  # empty launch + doorbell: the floor
    ctx.tune("doorbell", 0)  # the same with a stream synchronisation instead of the doorbell
    ping_sync_us = time_calls(lambda: check(L.pdsp_ctx_ping(ctx.h)), calls)
    fwd_sync_us = time_calls(lambda: check(L.pdsp_fft_forward_real(plan, vp(x), F64, 1, vp(ore), vp(oim))), calls)
    ctx.tune("doorbell", None)
    fwd_us = time_calls(lambda: check(L.pdsp_fft_forward_real(plan, vp(x), F64, 1, vp(ore), vp(oim))), calls)
    bins = n // 2 + 1
    amp, ph, pk = np.empty(bins), np.empty(bins), np.zeros(1, dtype=PEAK_F64)
    d1 = SpectrumDesc(sample_dtype=F64, frame_len=n, hop=n, batch=1, window=WINDOWS["hann"], sides=SIDES["one"],
                      sample_rate=48000.0, raw_magnitude=0)
    spec_us = time_calls(lambda: check(L.pdsp_spectrum(plan, C.byref(d1), vp(x), vp(amp), vp(ph), vp(pk))), calls)
    # the same two calls on the CPU port, one core
    po = oracle.FFT(n)
    cpu_fwd_us = time_calls(lambda: po.forward(x), calls)
    cpu_spec_us = time_calls(lambda: oracle.spectrum_batch(x[None, :], fftSize=n, sampleRate=48000.0, window="hann"), calls)
    rre, rim = po.forward(x)
    rel = float(np.linalg.norm((ore - rre.reshape(-1)) + 1j * (oim - rim.reshape(-1))) / np.linalg.norm(rre + 1j * rim))
    # crossover: frames per pdsp_spectrum call at which the GPU entry point beats the 1-core port
    cross = None
    per_frame_cpu = None
    xs = np.tile(x, (4096, 1))
    table = []
    for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 1024, 4096):
        a_b, p_b, k_b = np.empty((b, bins)), np.empty((b, bins)), np.zeros(b, dtype=PEAK_F64)
        db = SpectrumDesc(sample_dtype=F64, frame_len=n, hop=n, batch=b, window=WINDOWS["hann"], sides=SIDES["one"],
                          sample_rate=48000.0, raw_magnitude=0)
        g_us = time_calls(lambda: check(L.pdsp_spectrum(plan, C.byref(db), vp(xs), vp(a_b), vp(p_b), vp(k_b))), max(20, calls // 10))
        c_us = time_calls(lambda: oracle.spectrum_batch(xs[:b], fftSize=n, sampleRate=48000.0, window="hann"), max(5, calls // 40))
        table.append({"frames": b, "gpu_us": g_us, "cpu_1core_us": c_us})
        if cross is None and g_us < c_us:
            cross = b
        per_frame_cpu = c_us / b
    line = {
        "metric": "us_per_call", "value": fwd_us, "unit": "us", "n_gpus": 1, "steps": calls, "warmup": max(10, args.warmup * 4),
        "ms_per_step": fwd_us / 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": args.workload, "description": w["desc"], "fft_size": n},
        "latency": {"empty_launch_and_doorbell_us": ping_us, "empty_launch_and_stream_sync_us": ping_sync_us,
                    "Radix2Fft.forward_stream_sync_us": fwd_sync_us, "Radix2Fft.forward_us": fwd_us, "spectrum_us": spec_us, "cpu_port_forward_us": cpu_fwd_us,
                    "cpu_port_spectrum_us": cpu_spec_us, "gpu_overtakes_cpu_at_frames_per_call": cross,
                    "cpu_port_us_per_frame_batched": per_frame_cpu, "by_batch": table,
                    "api": "pdsp_fft_forward_real / pdsp_spectrum, pageable host arrays, one call = H2D + kernel + D2H + sync"},
        "cpu_baseline": {"value": cpu_fwd_us, "unit": "us", "cores": 1, "kind": "port",
                         "sample": "the same single frame through oracle/pragma_oracle.c (ctypes call overhead included on both arms)"},
        "e2e": {"value": fwd_us, "unit": "us", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 16 * n},
        "gpu_launches": ctx.launch_count, "parity": {"rel_l2_vs_oracle": rel, "bound": 1e-12 * np.log2(n)},
    }
    emit_line(line)
    ctx.close()
    return 0


class StdoutToStderr:
    """Keeps stdout clean for the single JSON line: anything libraries print to fd 1 meanwhile (NCCL's
    version banner, for one) goes to stderr; restore() gives the real stdout back for the result line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def restore(self):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="north_star", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = every rank owns the workload's frame count; strong = the frame count is split over the ranks")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the workload's)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: how per-frame peaks reach every rank - peer stores fused into the kernel epilogue (p2p) "
                         "or a side-stream NCCL all_gather")
    ap.add_argument("--comparator", action="store_true",
                    help="also time the same batch through torch.fft.rfft (cuFFT) + abs as context (amplitude-only workloads, 1 GPU)")
    ap.add_argument("--quick", action="store_true", help="profiling runs: skip the clock probe, CPU baseline, shorten e2e")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.frames:
        w["frames"] = args.frames
    global _STDOUT_GUARD
    _STDOUT_GUARD = StdoutToStderr()
    try:
        if args.impl == "reference":
            return run_reference(args, w)
        if w.get("kind") == "latency":
            return run_b200_latency(args, w)
        if w.get("kind") in ("c2c", "r2c_forward"):
            return run_b200_c2c(args, w)
        return run_b200(args, w)
    finally:
        if _STDOUT_GUARD is not None:
            _STDOUT_GUARD.restore()


if __name__ == "__main__":
    sys.exit(main())
