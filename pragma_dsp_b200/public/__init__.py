from .spectrum import spectrum, spectrum_batch  # noqa: F401
