"""C-ABI surface checks that need no GPU: the library loads, exports every symbol that
include/pragma_b200.h declares, the pure-host helpers match the oracle bit for bit, and the
product fails loudly (no CPU path) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pragma_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pdsp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pragma_dsp_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/pragma_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "python binding table and header disagree"
    assert L.pdsp_abi_version() == 1


def test_no_oracle_or_torch_in_product_package():
    """The product package must not import the oracle (or route compute through numpy/torch FFTs)."""
    pkg = os.path.join(ROOT, "pragma_dsp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f
                assert "np.fft" not in src and "torch.fft" not in src and "cufft" not in src.lower(), f


def test_host_helpers_match_oracle_bitwise():
    from pragma_dsp_b200.core import isPowerOfTwo, nextPowerOfTwo
    from pragma_dsp_b200.xform import binFrequencies, createWindow
    for n in (-5, 0, 1, 2, 3, 7, 8, 1000, 1024, 1025, 2**20 + 1):
        assert nextPowerOfTwo(n) == oracle.nextPowerOfTwo(n)
        assert isPowerOfTwo(n) == oracle.isPowerOfTwo(n)
    for t in ("rect", "hann", "hamming", "blackman"):
        for n in (1, 2, 8, 64, 1024, 4096):
            assert (createWindow(t, n) == oracle.createWindow(t, n)).all()
    for sides in ("one", "two"):
        assert (binFrequencies(1024, 48000.0, sides) == oracle.binFrequencies(1024, 48000.0, sides)).all()
    assert binFrequencies(1024, 48000.0)[-1] == 24000.0 and len(binFrequencies(7, 1.0)) == 4


def test_reference_error_messages():
    """Validation stays in the host layer with the reference's messages (SURVEY 8b Errors)."""
    from pragma_dsp_b200.xform import binFrequencies, createWindow
    with pytest.raises(ValueError, match="Window size must be positive, got 0"):
        createWindow("hann", 0)
    with pytest.raises(ValueError, match="Unsupported window type: kaiser"):
        createWindow("kaiser", 8)
    with pytest.raises(ValueError, match="FFT size must be positive, got 0"):
        binFrequencies(0, 1.0)
    with pytest.raises(ValueError, match="Sample rate must be positive, got -1"):
        binFrequencies(8, -1)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_fails_loudly_without_gpu():
    from pragma_dsp_b200 import PragmaB200Error, _lib, spectrum
    from pragma_dsp_b200.core import Radix2Fft
    h = C.c_void_p()
    assert _lib.lib().pdsp_ctx_create(0, C.byref(h)) != 0
    assert b"no CUDA device" in _lib.lib().pdsp_last_error() and b"no CPU path" in _lib.lib().pdsp_last_error()
    with pytest.raises(PragmaB200Error):
        spectrum(np.zeros(8))
    with pytest.raises(PragmaB200Error):
        Radix2Fft(8)
    with pytest.raises(ValueError, match="FFT size must be power of two, got 12"):
        Radix2Fft(12)  # validation precedes device work


def test_out_arrays_are_validated_before_native_code():
    """a caller-supplied `out` goes to native code as a raw pointer - short, strided, read-only or
    wrong-dtype planes must be refused by the host layer (no GPU needed: validation precedes device work)."""
    from pragma_dsp_b200.core import ComplexArray, _check_plane
    from pragma_dsp_b200.xform import fourier
    good = np.zeros(8)
    _check_plane(good, 8, "out")
    for bad, exc in ((np.zeros(7), ValueError), (np.zeros(16)[::2], ValueError), (np.zeros(8, dtype=np.float32), TypeError),
                     (np.zeros((2, 4)), ValueError), ([0.0] * 8, TypeError)):
        with pytest.raises(exc):
            _check_plane(bad, 8, "out")
    ro = np.zeros(8)
    ro.setflags(write=False)
    with pytest.raises(ValueError, match="writeable"):
        _check_plane(ro, 8, "out")
    c = ComplexArray(np.zeros(8), np.zeros(8))
    with pytest.raises(ValueError, match="length 8"):
        fourier.magnitude(c, np.zeros(4))
    with pytest.raises(ValueError, match="equal length"):
        fourier.phase(ComplexArray(np.zeros(8), np.zeros(4)))
    with pytest.raises(ValueError, match="length 8"):
        fourier.applyWindow(np.zeros(8), np.ones(8), np.zeros(9))
    with pytest.raises(ValueError, match="length 8"):
        fourier.fftShift(np.zeros(8), np.zeros(16)[::2][:8][:4])
    with pytest.raises(ValueError, match="equal length"):
        fourier.fftShiftComplex(ComplexArray(np.zeros(8), np.zeros(4)))
