"""Effect rung: mirrors /root/reference/src/effect/index.ts (FourierService, Fourier, FourierLive,
spectrumFx, spectrumStream) without the `effect` runtime, which has no Python equivalent:

* ``FourierLive()`` builds the service - two caches, ``fft(size)`` and ``window(type, size)``, whose
  returned instances are identity-stable like the reference's Maps (:30-48); used as a context
  manager it scopes the device plans the way ``Effect.provide(FourierLive)`` scopes the Layer.
* ``spectrumFx(samples, options)`` returns a function of the service (the Effect's "run").
* ``spectrumStream(frames, options)`` maps an iterable of frames to per-frame results, in order,
  1:1, empty in -> empty out (:190-194) - but groups ``chunk`` frames per launch so the stream is
  one batched kernel per chunk instead of one transform per fiber.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator

import numpy as np

from .. import _lib
from ..core import nextPowerOfTwo
from ..public.spectrum import spectrum_batch
from ..xform.fourier import FFT, createWindow


class FourierService:
    """src/effect/index.ts:17-20"""

    def __init__(self, context=None):
        self._ctx = context or _lib.default_context()
        self._fft: dict[int, FFT] = {}
        self._win: dict[str, np.ndarray] = {}

    def fft(self, size: int) -> FFT:
        cached = self._fft.get(size)
        if cached is not None:
            return cached
        created = FFT(size, context=self._ctx)
        self._fft[size] = created
        return created

    def window(self, type: str, size: int) -> np.ndarray:
        key = f"{type}:{size}"
        cached = self._win.get(key)
        if cached is not None:
            return cached
        created = createWindow(type, size)
        self._win[key] = created
        return created

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._fft.clear()
        self._win.clear()
        return False


Fourier = FourierService  # the Context.Tag ("pragma-dsp/Fourier", :22-25) is the class itself here


def FourierLive(context=None) -> FourierService:
    """src/effect/index.ts:27-51"""
    return FourierService(context)


def _one(service: FourierService, samples, options: dict):
    x = np.asarray(samples)
    size = options.get("fftSize")
    size = nextPowerOfTwo(x.shape[0]) if size is None else size
    service.fft(size)  # plan cached for the service's lifetime, like fftCache
    r = spectrum_batch(x.reshape(1, -1) if x.size else np.zeros((1, 0)), sampleRate=options.get("sampleRate", 1),
                       fftSize=size, window=options.get("window", "rect"), sides=options.get("sides", "one"),
                       precision=options.get("precision", "f64"), context=service._ctx)
    return _result(r, 0)


def _result(r, i):
    pk = r["peaks"][i]
    return {"frequencies": r["frequencies"], "amplitude": r["amplitude"][i], "phase": r["phase"][i],
            "peak": {"index": int(pk["index"]), "frequency": float(pk["frequency"]),
                     "amplitude": float(pk["amplitude"]), "phase": float(pk["phase"])}}


def spectrumFx(samples, options: dict | None = None) -> Callable[[FourierService], dict]:
    """src/effect/index.ts:181-188 - returns the effect; run it with a service: spectrumFx(x, o)(service)."""
    opts = dict(options or {})
    return lambda service: _one(service, samples, opts)


def spectrumStream(frames: Iterable, options: dict | None = None, *, service: FourierService | None = None,
                   chunk: int = 1024, lockstep: bool = False) -> Iterator[dict]:
    """src/effect/index.ts:190-194 - ordered 1:1 map over a stream of (Float32Array) frames.

    Batching contract.  Frames go through the ingestion ring (pdsp_ingest_*): each is copied into a pinned chunk as it
    arrives and results are yielded in arrival order, one per frame.  Chunks are cut ADAPTIVELY, not by a fixed size:

    * before the next frame is pulled from the source, every result that is already finished is yielded (a
      non-blocking poll, pdsp_ingest_ready);
    * whenever no chunk is in flight (the GPU is idle), the frames pushed so far are sent at once - a slow source
      (microphone, websocket) sees its result after one small launch, not after `chunk` frames;
    * while a chunk is in flight, arriving frames accumulate (up to `chunk`) and go out together when it completes - a
      fast source fills whole chunks and the copies overlap the kernels.

    So the latency of a result is bounded by one launch of at most `chunk` frames plus the time the source takes to
    yield the next frame.  A source that needs result i before it can produce frame i+1 (feedback) must pass
    lockstep=True: every frame is then sent and its result yielded before the next one is pulled - the reference's
    strictly sequential `Stream.mapEffect` behaviour, at one launch per frame.
    A frame of a different length or dtype drains the ring and opens a new one."""
    from ..public.ingest import IngestRing
    opts = dict(options or {})
    svc = service or FourierService()
    ring = None
    key = None
    pending = 0  # frames pushed and not yet yielded

    def take(r, n):
        """Yield n finished-or-in-flight frames (pop blocks only on chunks already sent)."""
        nonlocal pending
        while n > 0 and pending > 0:
            out = r.pop(min(n, pending, chunk))
            if out["count"] == 0:
                break
            for i in range(out["count"]):
                yield _result(out, i)
            pending -= out["count"]
            n -= out["count"]

    for frame in frames:
        f = np.asarray(frame)
        if f.dtype != np.float32 and f.dtype != np.float64:
            f = f.astype(np.float64)
        k = (f.shape, f.dtype)
        if ring is not None and k != key:
            ring.flush()
            yield from take(ring, pending)
            ring.close()
            ring = None
        if f.size == 0:  # spectrum([]) is a one-bin result; not worth a ring
            yield _one(svc, f, opts)
            continue
        if ring is None:
            size = opts.get("fftSize")
            size = nextPowerOfTwo(f.shape[0]) if size is None else size
            svc.fft(size)
            ring = IngestRing(f.shape[0], sampleRate=opts.get("sampleRate", 1), fftSize=size,
                              window=opts.get("window", "rect"), sides=opts.get("sides", "one"),
                              precision=opts.get("precision", "f64"), sample_dtype=f.dtype, framesPerChunk=chunk,
                              depth=3, context=svc._ctx)
            key = k
        while ring.push(f) == 0:  # ring full: hand back the oldest chunk, then retry
            yield from take(ring, 1)
        pending += 1
        if lockstep:
            ring.flush()
            yield from take(ring, pending)
            continue
        finished, in_flight, waiting = ring.ready()
        if in_flight == 0 and waiting > 0:  # nothing to overlap with: send what has arrived
            ring.flush()
        if finished:
            yield from take(ring, finished)
    if ring is not None:
        ring.flush()
        yield from take(ring, pending)
        ring.close()


__all__ = ["FourierService", "Fourier", "FourierLive", "spectrumFx", "spectrumStream"]
