"""ctypes binding of the CPU oracle (oracle/pragma_oracle.c).

TEST INFRASTRUCTURE ONLY - see the header of pragma_oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs; the product
package ``pragma_dsp_b200`` never imports this module.

The functions mirror the reference's names (src/core/fft.ts, src/xform/fourier.ts,
src/public/spectrum.ts) so the oracle-vs-golden tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpragma_oracle.so")

WINDOWS = {"rect": 0, "hann": 1, "hamming": 2, "blackman": 3}
SIDES = {"one": 0, "two": 1}


class Peak(C.Structure):
    _fields_ = [("index", C.c_int32), ("_pad", C.c_int32), ("frequency", C.c_double), ("amplitude", C.c_double),
                ("phase", C.c_double)]


PEAK_DTYPE = np.dtype([("index", "<i4"), ("_pad", "<i4"), ("frequency", "<f8"), ("amplitude", "<f8"), ("phase", "<f8")])


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "pragma_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.po_plan_create.restype = C.c_void_p
        L.po_plan_create.argtypes = [C.c_int]
        L.po_plan_destroy.argtypes = [C.c_void_p]
        L.po_fft_transform.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_int]
        L.po_create_window.argtypes = [C.c_int, C.c_int, dp]
        L.po_magnitude.argtypes = [dp, dp, C.c_int, dp]
        L.po_phase.argtypes = [dp, dp, C.c_int, dp]
        L.po_fft_shift.argtypes = [dp, C.c_int, dp]
        L.po_bin_frequencies.argtypes = [C.c_int, C.c_double, C.c_int, dp]
        L.po_spectrum_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
                                        C.c_longlong, C.c_double, C.c_int, dp, dp, C.c_void_p, C.c_int]
        L.po_spectrum_batch_lean.argtypes = L.po_spectrum_batch.argtypes
        L.po_fft_batch.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_longlong, C.c_int, C.c_int]
        L.po_bench_checksum.restype = C.c_double
        L.po_bench_checksum.argtypes = [dp, dp, C.c_int]
        L.po_next_power_of_two.argtypes = [C.c_int32]
        L.po_next_power_of_two.restype = C.c_int32
        L.po_is_power_of_two.argtypes = [C.c_int32]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def isPowerOfTwo(n: int) -> bool:
    return bool(lib().po_is_power_of_two(int(n)))


def nextPowerOfTwo(n: int) -> int:
    return int(lib().po_next_power_of_two(int(n)))


class Radix2Fft:
    """src/core/fft.ts:63-152, restated in C."""

    def __init__(self, size: int):
        h = lib().po_plan_create(int(size))
        if not h:
            raise ValueError(f"FFT size must be power of two, got {size}")
        self._h = h
        self.size = int(size)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.po_plan_destroy(h)
            self._h = None

    def _transform(self, re, im, inverse, threads=1):
        re = np.ascontiguousarray(re, dtype=np.float64)
        if re.shape[-1] != self.size:
            raise ValueError(f"FFT input length {re.shape[-1]} != size {self.size}")
        if im is not None:
            im = np.ascontiguousarray(im, dtype=np.float64)
            if im.shape != re.shape:
                raise ValueError(f"FFT input length {im.shape[-1]} != size {self.size}")
        out_re = np.empty_like(re)
        out_im = np.empty_like(re)
        batch = re.size // self.size
        lib().po_fft_batch(self._h, _dp(re), _dp(im), _dp(out_re), _dp(out_im), batch, int(inverse), int(threads))
        return out_re, out_im

    def forward(self, x, threads=1):
        return self._transform(x, None, False, threads)

    def forwardComplex(self, re, im, threads=1):
        return self._transform(re, im, False, threads)

    def inverse(self, re, im, threads=1):
        return self._transform(re, im, True, threads)


FFT = Radix2Fft  # src/xform/fourier.ts:69-96 is a pure delegate


def createWindow(type_: str, size: int) -> np.ndarray:
    if size <= 0:
        raise ValueError(f"Window size must be positive, got {size}")
    if type_ not in WINDOWS:
        raise ValueError(f"Unsupported window type: {type_}")
    out = np.empty(size, dtype=np.float64)
    rc = lib().po_create_window(WINDOWS[type_], int(size), _dp(out))
    assert rc == 0
    return out


def magnitude(re, im) -> np.ndarray:
    re = np.ascontiguousarray(re, dtype=np.float64)
    im = np.ascontiguousarray(im, dtype=np.float64)
    out = np.empty_like(re)
    lib().po_magnitude(_dp(re), _dp(im), re.size, _dp(out))
    return out


def phase(re, im) -> np.ndarray:
    re = np.ascontiguousarray(re, dtype=np.float64)
    im = np.ascontiguousarray(im, dtype=np.float64)
    out = np.empty_like(re)
    lib().po_phase(_dp(re), _dp(im), re.size, _dp(out))
    return out


def fftShift(x) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().po_fft_shift(_dp(x), x.size, _dp(out))
    return out


def binFrequencies(size: int, sampleRate: float, sides: str = "one") -> np.ndarray:
    if size <= 0:
        raise ValueError(f"FFT size must be positive, got {size}")
    if not sampleRate > 0:
        raise ValueError(f"Sample rate must be positive, got {sampleRate}")
    bins = size // 2 + 1 if sides == "one" else size
    out = np.empty(bins, dtype=np.float64)
    lib().po_bin_frequencies(int(size), float(sampleRate), SIDES[sides], _dp(out))
    return out


def spectrum_batch(samples, *, fftSize=None, frameLen=None, hop=None, batch=None, sampleRate=1.0, window="rect",
                   sides="one", want_amplitude=True, want_phase=True, want_peaks=True, threads=1, lean=False):
    """spectrum() (src/public/spectrum.ts:107-142) over a batch of frames.

    ``samples`` is a float32/float64 array; 2-D (batch, frameLen) or 1-D with explicit
    frameLen/hop/batch (STFT addressing).  Returns dict(amplitude, phase, peaks, frequencies).
    lean=True (bench.py's reference arm): Math.atan2 only where an output needs it (all bins for the phase array,
    the peak bin for a peak without one) instead of spectrum()'s unconditional phase pass.
    """
    samples = np.ascontiguousarray(samples)
    if samples.dtype not in (np.float32, np.float64):
        samples = samples.astype(np.float64)
    if samples.ndim == 2:
        batch = samples.shape[0] if batch is None else batch
        frameLen = samples.shape[1] if frameLen is None else frameLen
        hop = samples.shape[1] if hop is None else hop
    else:
        frameLen = samples.shape[0] if frameLen is None else frameLen
        hop = frameLen if hop is None else hop
        batch = 1 if batch is None else batch
    size = nextPowerOfTwo(frameLen) if fftSize is None else int(fftSize)
    plan = Radix2Fft(size)
    bins = size // 2 + 1 if sides == "one" else size
    amp = np.empty((batch, bins), dtype=np.float64) if want_amplitude else None
    ph = np.empty((batch, bins), dtype=np.float64) if want_phase else None
    peaks = np.zeros(batch, dtype=PEAK_DTYPE) if want_peaks else None
    fn = lib().po_spectrum_batch_lean if lean else lib().po_spectrum_batch
    used = fn(plan._h, WINDOWS[window], samples.ctypes.data_as(C.c_void_p),
                                   1 if samples.dtype == np.float64 else 0, int(frameLen), int(hop), int(batch),
                                   float(sampleRate), SIDES[sides], _dp(amp), _dp(ph),
                                   None if peaks is None else peaks.ctypes.data_as(C.c_void_p), int(threads))
    return {"amplitude": amp, "phase": ph, "peaks": peaks, "frequencies": binFrequencies(size, sampleRate, sides),
            "threads": used}


def spectrum(samples, sampleRate=1.0, fftSize=None, window="rect", sides="one"):
    """Single-frame spectrum(); returns dict(frequencies, amplitude, phase, peak)."""
    x = np.asarray(samples)
    if x.dtype != np.float32:
        x = x.astype(np.float64)
    size = nextPowerOfTwo(x.shape[0]) if fftSize is None else int(fftSize)
    if x.shape[0] == 0:
        x = np.zeros(1, dtype=np.float64)
        r = spectrum_batch(x, fftSize=size, frameLen=0, hop=0, batch=1, sampleRate=sampleRate, window=window,
                           sides=sides)
    else:
        r = spectrum_batch(x[None, :], fftSize=size, sampleRate=sampleRate, window=window, sides=sides)
    pk = r["peaks"][0]
    return {"frequencies": r["frequencies"], "amplitude": r["amplitude"][0], "phase": r["phase"][0],
            "peak": {"index": int(pk["index"]), "frequency": float(pk["frequency"]),
                     "amplitude": float(pk["amplitude"]), "phase": float(pk["phase"])}}


def bench_checksum(re, im) -> float:
    re = np.ascontiguousarray(re, dtype=np.float64)
    im = np.ascontiguousarray(im, dtype=np.float64)
    return float(lib().po_bench_checksum(_dp(re), _dp(im), re.size))


def max_threads() -> int:
    return int(lib().po_max_threads())
