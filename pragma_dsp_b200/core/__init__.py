"""Expert rung: mirrors /root/reference/src/core/fft.ts (ComplexArray, createComplexArray,
isPowerOfTwo, nextPowerOfTwo, Radix2Fft) on top of the C-ABI."""
from __future__ import annotations

import numpy as np

from .. import _lib
from .._lib import F32, F64, check, lib, ptr


class ComplexArray:
    """src/core/fft.ts:1-4 - split (planar) complex container of two float64 arrays."""
    __slots__ = ("real", "imag")

    def __init__(self, real: np.ndarray, imag: np.ndarray):
        self.real = real
        self.imag = imag


def createComplexArray(size: int, fill: float = 0) -> ComplexArray:
    """src/core/fft.ts:6-14"""
    real = np.zeros(int(size), dtype=np.float64)
    imag = np.zeros(int(size), dtype=np.float64)
    if fill != 0:
        real.fill(fill)
        imag.fill(fill)
    return ComplexArray(real, imag)


def isPowerOfTwo(n: int) -> bool:
    """src/core/fft.ts:16"""
    return bool(lib().pdsp_is_power_of_two(int(n))) if -2**31 <= n < 2**31 else False


def nextPowerOfTwo(n: int) -> int:
    """src/core/fft.ts:18-23"""
    return int(lib().pdsp_next_power_of_two(int(n)))


def _check_plane(a, size: int, what: str) -> None:
    """A caller-supplied output plane is handed to native code as a raw pointer: it must be exactly what the C ABI
    writes - 1-D, `size` float64 elements, C-contiguous, writeable.  (The reference silently drops out-of-range
    typed-array writes; a native memcpy would corrupt the heap instead, so this layer refuses.)"""
    if not isinstance(a, np.ndarray):
        raise TypeError(f"{what} must be a numpy float64 array")
    if a.dtype != np.float64:
        raise TypeError(f"{what} must be float64 (Float64Array), got {a.dtype}")
    if a.ndim != 1 or a.shape[0] != size:
        raise ValueError(f"{what} must have length {size}, got shape {a.shape}")
    if not a.flags.c_contiguous:
        raise ValueError(f"{what} must be contiguous")
    if not a.flags.writeable:
        raise ValueError(f"{what} must be writeable")


def _as_samples(x):
    """ArrayLike<number> -> contiguous float32/float64 (`?? 0` for missing elements is a JS-only notion)."""
    a = np.asarray(x)
    if a.dtype == np.float32 or a.dtype == np.float64:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a, dtype=np.float64)


class Radix2Fft:
    """src/core/fft.ts:63-152.  `size` is checked here so the message is the reference's."""

    def __init__(self, size: int, *, context: "_lib.Context | None" = None):
        if not isPowerOfTwo(size):
            raise ValueError(f"FFT size must be power of two, got {size}")
        self.size = int(size)
        self._ctx = context or _lib.default_context()
        self._plan = self._ctx.plan(self.size, F64)

    def _out(self, out):
        if out is None:
            return createComplexArray(self.size)
        _check_plane(out.real, self.size, "out.real")
        _check_plane(out.imag, self.size, "out.imag")
        return out

    def forward(self, input, out: ComplexArray | None = None) -> ComplexArray:
        """:77-79 - real input of length size -> all N bins; returns the same `out` object."""
        x = _as_samples(input)
        if x.ndim != 1 or x.shape[0] != self.size:
            raise ValueError(f"FFT input length {x.shape[0] if x.ndim else 0} != size {self.size}")
        if self.size > 16384 and x.dtype != np.float64:
            x = x.astype(np.float64)  # the multi-pass path reads planes in the plan's precision
        result = self._out(out)
        check(lib().pdsp_fft_forward_real(self._plan, ptr(x), F64 if x.dtype == np.float64 else F32, 1,
                                          ptr(result.real), ptr(result.imag)))
        return result

    def _complex(self, input: ComplexArray, out, fn):
        re = np.ascontiguousarray(input.real, dtype=np.float64)
        im = np.ascontiguousarray(input.imag, dtype=np.float64)
        if re.ndim != 1 or re.shape[0] != self.size:
            raise ValueError(f"FFT input length {re.shape[0] if re.ndim else 0} != size {self.size}")
        if im.ndim != 1 or im.shape[0] != self.size:
            raise ValueError(f"FFT input length {im.shape[0] if im.ndim else 0} != size {self.size}")
        result = self._out(out)
        check(fn(self._plan, ptr(re), ptr(im), 1, ptr(result.real), ptr(result.imag)))
        return result

    def forwardComplex(self, input: ComplexArray, out: ComplexArray | None = None) -> ComplexArray:
        """:81-83"""
        return self._complex(input, out, lib().pdsp_fft_forward_complex)

    def inverse(self, input: ComplexArray, out: ComplexArray | None = None) -> ComplexArray:
        """:85-87"""
        return self._complex(input, out, lib().pdsp_fft_inverse)

    # ---- additive batched forms (frames are rows)
    def forward_batch(self, frames) -> tuple[np.ndarray, np.ndarray]:
        x = _as_samples(frames)
        if x.ndim != 2 or x.shape[1] != self.size:
            raise ValueError(f"FFT input length {x.shape[-1]} != size {self.size}")
        re = np.empty(x.shape, dtype=np.float64)
        im = np.empty(x.shape, dtype=np.float64)
        check(lib().pdsp_fft_forward_real(self._plan, ptr(x), F64 if x.dtype == np.float64 else F32, x.shape[0],
                                          ptr(re), ptr(im)))
        return re, im

    def complex_batch(self, re, im, inverse=False) -> tuple[np.ndarray, np.ndarray]:
        re = np.ascontiguousarray(re, dtype=np.float64)
        im = np.ascontiguousarray(im, dtype=np.float64)
        if re.ndim != 2 or re.shape[1] != self.size or im.shape != re.shape:
            raise ValueError(f"FFT input length {re.shape[-1]} != size {self.size}")
        ore, oim = np.empty_like(re), np.empty_like(re)
        fn = lib().pdsp_fft_inverse if inverse else lib().pdsp_fft_forward_complex
        check(fn(self._plan, ptr(re), ptr(im), re.shape[0], ptr(ore), ptr(oim)))
        return ore, oim


__all__ = ["ComplexArray", "createComplexArray", "isPowerOfTwo", "nextPowerOfTwo", "Radix2Fft"]
