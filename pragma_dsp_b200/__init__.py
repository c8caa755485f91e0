"""pragma_dsp_b200 - B200 (sm_100a) implementation of pragma-dsp's FFT/spectrum hot path.

Same ladder as the reference's subpath exports (package.json:14-55):
    pragma_dsp_b200            -> spectrum                      (src/index.ts)
    pragma_dsp_b200.core       -> Radix2Fft, createComplexArray (src/core)
    pragma_dsp_b200.xform      -> FFT, createWindow, magnitude, phase, binFrequencies (src/xform/fourier)
    pragma_dsp_b200.effect     -> FourierLive, spectrumFx, spectrumStream (src/effect)
All compute goes through libpragma_b200.so (include/pragma_b200.h); there is no CPU fallback.
"""
from ._lib import PragmaB200Error  # noqa: F401
from .group import DeviceGroup  # noqa: F401
from .public.ingest import IngestRing  # noqa: F401
from .public.spectrum import spectrum, spectrum_batch, stft  # noqa: F401

__all__ = ["spectrum", "spectrum_batch", "stft", "IngestRing", "DeviceGroup", "PragmaB200Error"]
