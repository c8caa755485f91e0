"""Power rung: mirrors /root/reference/src/xform/fourier.ts (FFT, createWindow, magnitude, phase,
binFrequencies) on top of the C-ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import SIDES, WINDOWS, check, lib, ptr
from ..core import ComplexArray, Radix2Fft, _check_plane, createComplexArray, isPowerOfTwo

WindowType = str  # "rect" | "hann" | "hamming" | "blackman"
FftSides = str    # "one" | "two"


def createWindow(type: WindowType, size: int) -> np.ndarray:
    """src/xform/fourier.ts:14-52 - symmetric window as a fresh float64 array."""
    if size <= 0:
        raise ValueError(f"Window size must be positive, got {size}")
    if type not in WINDOWS:
        raise ValueError(f"Unsupported window type: {type}")
    out = np.empty(int(size), dtype=np.float64)
    check(lib().pdsp_create_window(WINDOWS[type], int(size), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


class FFT:
    """src/xform/fourier.ts:69-96 - facade over Radix2Fft."""

    def __init__(self, size: int, *, context: "_lib.Context | None" = None):
        if not isPowerOfTwo(size):
            raise ValueError(f"FFT size must be power of two, got {size}")
        self.size = int(size)
        self._kernel = Radix2Fft(size, context=context)

    def forward(self, input, out: ComplexArray | None = None) -> ComplexArray:
        return self._kernel.forward(input, out)

    def forwardComplex(self, input: ComplexArray, out: ComplexArray | None = None) -> ComplexArray:
        return self._kernel.forwardComplex(input, out)

    def inverse(self, input: ComplexArray, out: ComplexArray | None = None) -> ComplexArray:
        return self._kernel.inverse(input, out)

    def createComplexArray(self, fill: float = 0) -> ComplexArray:
        return createComplexArray(self.size, fill)

    def forward_batch(self, frames):
        return self._kernel.forward_batch(frames)

    def complex_batch(self, re, im, inverse=False):
        return self._kernel.complex_batch(re, im, inverse)


def _elementwise(fn, input: ComplexArray, out):
    re = np.ascontiguousarray(input.real, dtype=np.float64)
    im = np.ascontiguousarray(input.imag, dtype=np.float64)
    if re.ndim != 1 or im.shape != re.shape:
        # the reference reads a missing imaginary element as `?? 0`; a native read past a shorter plane is not an option
        raise ValueError(f"real and imaginary planes must be 1-D and of equal length, got {re.shape} and {im.shape}")
    if out is not None:
        _check_plane(out, re.shape[0], "out")
    result = out if out is not None else np.empty(re.shape[0], dtype=np.float64)
    check(fn(_lib.default_context().h, ptr(re), ptr(im), re.shape[0], ptr(result)))
    return result


def magnitude(input: ComplexArray, out: np.ndarray | None = None) -> np.ndarray:
    """src/xform/fourier.ts:98-109 - hypot(re, im) for all bins."""
    return _elementwise(lib().pdsp_magnitude, input, out)


def phase(input: ComplexArray, out: np.ndarray | None = None) -> np.ndarray:
    """src/xform/fourier.ts:111-120 - atan2(im, re) for all bins."""
    return _elementwise(lib().pdsp_phase, input, out)


def applyWindow(input, window, out: np.ndarray | None = None) -> np.ndarray:
    """src/xform/fourier.ts:54-67 - elementwise product (the fused spectrum path never materialises it)."""
    x = np.ascontiguousarray(input, dtype=np.float64)
    w = np.ascontiguousarray(window, dtype=np.float64)
    if x.ndim != 1 or w.ndim != 1 or x.shape[0] != w.shape[0]:
        raise ValueError("Window length must match input length.")
    if out is not None:
        _check_plane(out, x.shape[0], "out")
    result = out if out is not None else np.empty(x.shape[0], dtype=np.float64)
    check(lib().pdsp_apply_window(_lib.default_context().h, ptr(x), ptr(w), x.shape[0], ptr(result)))
    return result


def fftShift(input, out: np.ndarray | None = None) -> np.ndarray:
    """src/xform/fourier.ts:122-134 - rotate by floor(n/2) so DC sits in the middle."""
    x = np.ascontiguousarray(input, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError(f"fftShift expects a 1-D array, got shape {x.shape}")
    if out is not None:
        _check_plane(out, x.shape[0], "out")
    result = out if out is not None else np.empty(x.shape[0], dtype=np.float64)
    if x.shape[0]:
        check(lib().pdsp_fft_shift(_lib.default_context().h, ptr(x), x.shape[0], ptr(result)))
    return result


def fftShiftComplex(input: ComplexArray, out: ComplexArray | None = None) -> ComplexArray:
    """src/xform/fourier.ts:136-145"""
    if np.shape(input.imag) != np.shape(input.real):
        raise ValueError("real and imaginary planes must be of equal length")
    result = out if out is not None else createComplexArray(input.real.shape[0])
    fftShift(input.real, result.real)
    fftShift(input.imag, result.imag)
    return result


def binFrequencies(size: int, sampleRate: float, sides: FftSides = "one") -> np.ndarray:
    """src/xform/fourier.ts:147-165"""
    if size <= 0:
        raise ValueError(f"FFT size must be positive, got {size}")
    if not sampleRate > 0:
        raise ValueError(f"Sample rate must be positive, got {sampleRate}")
    bins = size // 2 + 1 if sides == "one" else size
    out = np.empty(bins, dtype=np.float64)
    check(lib().pdsp_bin_frequencies(int(size), float(sampleRate), SIDES[sides],
                                     out.ctypes.data_as(C.POINTER(C.c_double)), None))
    return out


__all__ = ["FFT", "createWindow", "applyWindow", "magnitude", "phase", "fftShift", "fftShiftComplex", "binFrequencies",
           "WindowType", "FftSides"]
