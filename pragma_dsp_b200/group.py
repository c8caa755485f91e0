"""Device groups: the frame-sharded multi-GPU forms of spectrum() (BASELINE config C5).

Thin wrapper over pdsp_group_* (include/pragma_b200.h).  The reference has no counterpart - it is single-threaded;
frames are independent (spectrumStream is a pure 1:1 map, /root/reference/src/effect/index.ts:190-194), so device i of
n owns the contiguous block [i*ceil(F/n), (i+1)*ceil(F/n)) of the frame range.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import F32, F64, PEAK_F32, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib
from .core import _as_samples, isPowerOfTwo, nextPowerOfTwo
from .xform.fourier import binFrequencies


def block_of(batch: int, n: int, i: int) -> tuple[int, int]:
    """(first frame, frame count) of device i's block - the partition pdsp_group_* uses."""
    per = -(-batch // n)
    a, b = min(batch, per * i), min(batch, per * (i + 1))
    return a, b - a


class DeviceGroup:
    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        self._h = C.c_void_p()
        check(lib().pdsp_group_create(devs, len(devices), C.byref(self._h)))
        self.devices = [int(d) for d in devices]

    def __len__(self):
        return len(self.devices)

    def ctx_handle(self, i: int) -> C.c_void_p:
        return C.c_void_p(lib().pdsp_group_ctx(self._h, int(i)))

    def spectrum_batch(self, samples, *, sampleRate: float = 1, fftSize: int | None = None, window: str = "rect",
                       sides: str = "one", frameLen: int | None = None, hop: int | None = None, batch: int | None = None,
                       precision="f64", outputs=("amplitude", "phase", "peak")):
        """pragma_dsp_b200.spectrum_batch over all devices of the group: host frames in, host results out."""
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
        prec = F64 if precision in (F64, "f64", "float64", None) else F32
        dt = np.float64 if prec == F64 else np.float32
        bins = size // 2 + 1 if sides == "one" else size
        amp = np.empty((batch, bins), dtype=dt) if "amplitude" in outputs else None
        ph = np.empty((batch, bins), dtype=dt) if "phase" in outputs else None
        peaks = np.zeros(batch, dtype=PEAK_F64 if prec == F64 else PEAK_F32) if "peak" in outputs else None
        d = SpectrumDesc(sample_dtype=F64 if x.dtype == np.float64 else F32, frame_len=int(frameLen), hop=int(hop), batch=int(batch),
                         window=WINDOWS[window], sides=SIDES[sides], sample_rate=float(sampleRate), raw_magnitude=0, fft_shift=0)
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)  # noqa: E731
        check(lib().pdsp_group_spectrum(self._h, size, prec, C.byref(d), p(x) if x.size else None, p(amp), p(ph), p(peaks)))
        return {"frequencies": binFrequencies(size, sampleRate, sides), "amplitude": amp, "phase": ph, "peaks": peaks}

    def spectrum_dev(self, size: int, precision: int, desc: SpectrumDesc, d_samples, d_amplitude=None, d_phase=None, d_peaks=None,
                     gather_root: int = -1, d_amplitude_all=None, d_phase_all=None) -> None:
        """pdsp_group_spectrum_dev: lists of device pointers (ints), one per device; see include/pragma_b200.h."""
        n = len(self.devices)
        arr = lambda lst: None if lst is None else (C.c_void_p * n)(*[None if v is None else int(v) for v in lst])  # noqa: E731
        vp = lambda v: None if v is None else C.c_void_p(int(v))  # noqa: E731
        check(lib().pdsp_group_spectrum_dev(self._h, int(size), int(precision), C.byref(desc), arr(d_samples), arr(d_amplitude),
                                            arr(d_phase), arr(d_peaks), int(gather_root), vp(d_amplitude_all), vp(d_phase_all)))

    def sync(self) -> None:
        check(lib().pdsp_group_sync(self._h))

    def close(self) -> None:
        if self._h:
            check(lib().pdsp_group_destroy(self._h))
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
