from .spectrum import spectrum, spectrum_batch, stft  # noqa: F401
from .ingest import IngestRing  # noqa: F401
