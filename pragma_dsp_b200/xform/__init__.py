from .fourier import (FFT, applyWindow, binFrequencies, createWindow, fftShift, fftShiftComplex, magnitude,  # noqa: F401
                      phase)
