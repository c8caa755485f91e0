from .fourier import FFT, binFrequencies, createWindow, magnitude, phase  # noqa: F401
