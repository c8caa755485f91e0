"""Beginner rung: mirrors /root/reference/src/public/spectrum.ts - `spectrum(samples, options)`.

One call = one fused kernel launch (frame build, window, FFT, magnitude, phase, scaling, peak).
`spectrum_batch` is the additive batched form used by the Effect layer and the benchmarks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import F32, F64, PEAK_F32, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib, ptr
from ..core import _as_samples, isPowerOfTwo, nextPowerOfTwo
from ..xform.fourier import binFrequencies


def _precision(p):
    if p in (F64, "f64", "float64", None):
        return F64
    if p in (F32, "f32", "float32"):
        return F32
    raise ValueError(f"unknown precision {p!r}")


def spectrum_batch(samples, *, sampleRate: float = 1, fftSize: int | None = None, window: str = "rect",
                   sides: str = "one", frameLen: int | None = None, hop: int | None = None, batch: int | None = None,
                   precision="f64", outputs=("amplitude", "phase", "peak"), raw_magnitude=False, shift=False,
                   context=None):
    """Batched spectrum(): `samples` is (batch, frameLen), or 1-D with frameLen/hop/batch (STFT view).

    Returns dict(frequencies, amplitude, phase, peaks) with arrays in the plan precision.  shift=True (two-sided
    only) stores the amplitude / phase rows fftShift-ed - fftShift(amplitude) of src/xform/fourier.ts:122-134 fused
    into the kernel's stores; `frequencies` is left as binFrequencies() returns it and peak.index is the unshifted bin."""
    ctx = context or _lib.default_context()
    x = _as_samples(samples)
    if x.ndim == 2:
        batch = x.shape[0] if batch is None else batch
        frameLen = x.shape[1] if frameLen is None else frameLen
        hop = x.shape[1] if hop is None else hop
    else:
        frameLen = x.shape[0] if frameLen is None else frameLen
        hop = frameLen if hop is None else hop
        batch = 1 if batch is None else batch
    if batch > 0 and frameLen > 0 and (batch - 1) * hop + frameLen > x.size:
        raise ValueError("frames exceed the samples buffer")
    size = nextPowerOfTwo(frameLen) if fftSize is None else int(fftSize)
    if not isPowerOfTwo(size):
        raise ValueError(f"FFT size must be power of two, got {size}")
    if window not in WINDOWS:
        raise ValueError(f"Unsupported window type: {window}")
    if not sampleRate > 0:
        raise ValueError(f"Sample rate must be positive, got {sampleRate}")
    prec = _precision(precision)
    plan = ctx.plan(size, prec)
    dt = np.float64 if prec == F64 else np.float32
    bins = size // 2 + 1 if sides == "one" else size
    amp = np.empty((batch, bins), dtype=dt) if "amplitude" in outputs else None
    ph = np.empty((batch, bins), dtype=dt) if "phase" in outputs else None
    peaks = np.zeros(batch, dtype=PEAK_F64 if prec == F64 else PEAK_F32) if "peak" in outputs else None
    d = SpectrumDesc(sample_dtype=F64 if x.dtype == np.float64 else F32, frame_len=int(frameLen), hop=int(hop),
                     batch=int(batch), window=WINDOWS[window], sides=SIDES[sides], sample_rate=float(sampleRate),
                     raw_magnitude=int(bool(raw_magnitude)), fft_shift=int(bool(shift)))
    check(lib().pdsp_spectrum(plan, C.byref(d), ptr(x) if x.size else None, ptr(amp), ptr(ph), ptr(peaks)))
    return {"frequencies": binFrequencies(size, sampleRate, sides), "amplitude": amp, "phase": ph, "peaks": peaks}


def spectrum(samples, options: dict | None = None, **kw):
    """src/public/spectrum.ts:107-142.  Options: sampleRate=1, fftSize=nextPowerOfTwo(len), window="rect",
    sides="one" (+ optional precision="f64").  Returns dict(frequencies, amplitude, phase, peak)."""
    opts = dict(options or {})
    opts.update(kw)
    x = _as_samples(samples)
    n = x.shape[0]
    size = opts.get("fftSize")
    size = nextPowerOfTwo(n) if size is None else size
    r = spectrum_batch(x.reshape(1, n) if n else np.zeros((1, 0)), sampleRate=opts.get("sampleRate", 1), fftSize=size,
                       window=opts.get("window", "rect"), sides=opts.get("sides", "one"),
                       precision=opts.get("precision", "f64"), context=opts.get("context"))
    pk = r["peaks"][0]
    return {"frequencies": r["frequencies"], "amplitude": r["amplitude"][0], "phase": r["phase"][0],
            "peak": {"index": int(pk["index"]), "frequency": float(pk["frequency"]),
                     "amplitude": float(pk["amplitude"]), "phase": float(pk["phase"])}}


def stft(signal, *, fftSize: int, hopSize: int, window: str = "hann", sampleRate: float = 1, sides: str = "one",
         precision="f64", outputs=("amplitude",), raw_magnitude=False, context=None):
    """STFT / spectrogram (the reference's v0.2 roadmap item, ROADMAP.md:31-45; BASELINE config C3 does it
    through spectrumStream): frames are overlapping views signal[f*hopSize : f*hopSize + fftSize]; no frame
    matrix is materialised - the hop is an address stride inside the fused kernel.

    Returns dict(frequencies, times, amplitude (frames x bins), phase, peaks) in the plan precision."""
    x = _as_samples(signal)
    if x.ndim != 1:
        raise ValueError("stft expects a 1-D signal")
    if hopSize <= 0:
        raise ValueError(f"hop size must be positive, got {hopSize}")
    frames = 0 if x.shape[0] < fftSize else (x.shape[0] - fftSize) // hopSize + 1
    r = spectrum_batch(x, sampleRate=sampleRate, fftSize=fftSize, window=window, sides=sides, frameLen=fftSize,
                       hop=hopSize, batch=frames, precision=precision, outputs=outputs, raw_magnitude=raw_magnitude,
                       context=context)
    r["times"] = np.arange(frames, dtype=np.float64) * (hopSize / float(sampleRate))
    return r
