"""Ingestion ring: frames pushed one at a time (or in blocks), results popped in arrival order.

Thin wrapper over pdsp_ingest_* (include/pragma_b200.h).  The reference's streaming entry point is
spectrumStream (/root/reference/src/effect/index.ts:190-194): an ordered 1:1 map over Stream<Float32Array>;
the ring is what keeps the GPU fed from such a source without the caller assembling batches.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import F32, F64, PEAK_F32, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib
from ..core import isPowerOfTwo, nextPowerOfTwo
from ..xform.fourier import binFrequencies


class IngestRing:
    def __init__(self, frameLen: int, *, sampleRate: float = 1, fftSize: int | None = None, window: str = "rect",
                 sides: str = "one", precision="f64", sample_dtype=np.float32, outputs=("amplitude", "phase", "peak"),
                 framesPerChunk: int = 4096, depth: int = 3, context=None):
        ctx = context or _lib.default_context()
        size = nextPowerOfTwo(frameLen) if fftSize is None else int(fftSize)
        if not isPowerOfTwo(size):
            raise ValueError(f"FFT size must be power of two, got {size}")
        if window not in WINDOWS:
            raise ValueError(f"Unsupported window type: {window}")
        if not sampleRate > 0:
            raise ValueError(f"Sample rate must be positive, got {sampleRate}")
        self.frameLen, self.size, self.sides = int(frameLen), size, sides
        self.sample_dtype = np.dtype(sample_dtype)
        if self.sample_dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("frames must be float32 or float64")
        prec = F64 if precision in ("f64", F64, np.float64) else F32
        self._dt = np.float64 if prec == F64 else np.float32
        self._pk = PEAK_F64 if prec == F64 else PEAK_F32
        self.bins = size // 2 + 1 if sides == "one" else size
        self.outputs = tuple(outputs)
        self.frequencies = binFrequencies(size, sampleRate, sides)
        plan = ctx.plan(size, prec)
        d = SpectrumDesc(sample_dtype=F64 if self.sample_dtype == np.float64 else F32, frame_len=self.frameLen,
                         hop=self.frameLen, batch=0, window=WINDOWS[window], sides=SIDES[sides],
                         sample_rate=float(sampleRate), raw_magnitude=0, fft_shift=0)
        self._h = C.c_void_p()
        self._ctx = ctx  # keeps the context (and its plans) alive for the ring's lifetime
        check(lib().pdsp_ingest_open(plan, C.byref(d), int("amplitude" in outputs), int("phase" in outputs),
                                     int("peak" in outputs), int(framesPerChunk), int(depth), C.byref(self._h)))

    def push(self, frames) -> int:
        """Copies frames ((k, frameLen) or (frameLen,)) into the ring; returns how many were accepted (fewer than k:
        the ring is full - pop() first)."""
        x = np.ascontiguousarray(frames, dtype=self.sample_dtype)
        x = x.reshape(-1, self.frameLen)
        acc = C.c_int64(0)
        check(lib().pdsp_ingest_push(self._h, x.ctypes.data_as(C.c_void_p), x.shape[0], 0, C.byref(acc)))
        return int(acc.value)

    def push_pinned(self, frames) -> int:
        """As push() for frames that live in pinned host memory (pdsp_host_alloc): no host copy - the device reads them
        by DMA after this call returns, so they must stay unchanged until popped."""
        x = np.asarray(frames)
        if x.dtype != self.sample_dtype or not x.flags.c_contiguous:
            raise ValueError("push_pinned needs a contiguous array of the ring's sample dtype (it is not copied)")
        x = x.reshape(-1, self.frameLen)
        acc = C.c_int64(0)
        check(lib().pdsp_ingest_push_pinned(self._h, x.ctypes.data_as(C.c_void_p), x.shape[0], 0, C.byref(acc)))
        return int(acc.value)

    def ready(self) -> tuple[int, int, int]:
        """(finished, in_flight, pending) frames, without blocking: what pop() would return at once, what was sent and
        is still being processed, what sits in the partially filled chunk."""
        fin, fly, pend = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(lib().pdsp_ingest_ready(self._h, C.byref(fin), C.byref(fly), C.byref(pend)))
        return int(fin.value), int(fly.value), int(pend.value)

    def flush(self) -> None:
        check(lib().pdsp_ingest_flush(self._h))

    def pop(self, max_frames: int) -> dict:
        """Up to max_frames finished frames, in arrival order (waits for chunks already sent)."""
        amp = np.empty((max_frames, self.bins), dtype=self._dt) if "amplitude" in self.outputs else None
        ph = np.empty((max_frames, self.bins), dtype=self._dt) if "phase" in self.outputs else None
        pk = np.zeros(max_frames, dtype=self._pk) if "peak" in self.outputs else None
        got = C.c_int64(0)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(lib().pdsp_ingest_pop(self._h, p(amp), p(ph), p(pk), int(max_frames), C.byref(got)))
        k = int(got.value)
        return {"frequencies": self.frequencies, "amplitude": None if amp is None else amp[:k],
                "phase": None if ph is None else ph[:k], "peaks": None if pk is None else pk[:k], "count": k}

    def pop_into(self, amplitude=None, phase=None, peaks=None, max_frames: int | None = None) -> int:
        """pop() into caller-owned arrays (rows appended densely from row 0): no allocation per call - what a steady
        stream should use.  Arrays must be C-contiguous, of the plan precision / peak dtype, with room for max_frames."""
        cap = []
        for a, dt, name in ((amplitude, self._dt, "amplitude"), (phase, self._dt, "phase"), (peaks, self._pk, "peaks")):
            if a is None:
                continue
            if a.dtype != dt or not a.flags.c_contiguous or not a.flags.writeable:
                raise ValueError(f"{name} must be a writeable contiguous array of dtype {np.dtype(dt)}")
            cap.append(a.shape[0])
        if not cap:
            raise ValueError("no output array given")
        n = min(cap) if max_frames is None else int(max_frames)
        if n > min(cap):
            raise ValueError("max_frames exceeds the room in the output arrays")
        got = C.c_int64(0)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(lib().pdsp_ingest_pop(self._h, p(amplitude), p(phase), p(peaks), n, C.byref(got)))
        return int(got.value)

    def close(self) -> None:
        if self._h:
            check(lib().pdsp_ingest_close(self._h))
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
