from .spectrum import spectrum, spectrum_batch, stft  # noqa: F401
