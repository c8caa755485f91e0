"""The reference's own test suite (tests/reference_suite.py) replayed on the CUDA path through the
reference-shaped host API -> C-ABI -> sm_100a kernels.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import reference_suite as suite

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def impl():
    from pragma_dsp_b200 import spectrum
    from pragma_dsp_b200.core import ComplexArray
    from pragma_dsp_b200 import xform
    from pragma_dsp_b200.xform import FFT, magnitude, phase

    class B200Impl:
        @staticmethod
        def forward(n, x):
            r = FFT(n).forward(x)
            return r.real, r.imag

        @staticmethod
        def inverse(n, re, im):
            r = FFT(n).inverse(ComplexArray(np.asarray(re), np.asarray(im)))
            return r.real, r.imag

        @staticmethod
        def magnitude(re, im):
            return magnitude(ComplexArray(np.asarray(re), np.asarray(im)))

        @staticmethod
        def phase(re, im):
            return phase(ComplexArray(np.asarray(re), np.asarray(im)))

        createWindow = staticmethod(xform.createWindow)
        binFrequencies = staticmethod(xform.binFrequencies)

        @staticmethod
        def spectrum(x, sampleRate=1.0, fftSize=None, window="rect", sides="one"):
            return spectrum(x, {"sampleRate": sampleRate, "fftSize": fftSize, "window": window, "sides": sides})

    return B200Impl


@pytest.mark.parametrize("check", suite.ALL_CHECKS, ids=lambda f: f.__name__)
def test_b200_replays_reference_suite(impl, check):
    check(impl)


def test_b200_bench_run_checksums(impl):
    """bench/run.ts:20-35 - the guardrail checksums printed to 6 decimals."""
    import oracle
    suite.check_bench_checksums(impl, oracle.bench_checksum)


def test_out_identity_and_fill():
    """forward(input, out) returns the same object and fills the caller's planes
    (src/core/fft.ts:106,150; test/fluent/chain.test.ts:68-76)."""
    from pragma_dsp_b200.core import Radix2Fft, createComplexArray
    fft = Radix2Fft(8)
    out = createComplexArray(8, 7.0)
    r = fft.forward([1, 0, 0, 0, 0, 0, 0, 0], out)
    assert r is out
    assert (out.real == 1).all() and (out.imag == 0).all()
    back = fft.inverse(out)
    assert np.abs(back.real - np.array([1, 0, 0, 0, 0, 0, 0, 0])).max() == 0


def test_effect_layer_parity():
    """test/reallife/effect.test.ts - spectrumFx == spectrum bit for bit, cache identity, stream order/count."""
    from conftest import reallife
    from pragma_dsp_b200 import spectrum
    from pragma_dsp_b200.effect import FourierLive, spectrumFx, spectrumStream
    cases = reallife("pure_sine")[:3]
    with FourierLive() as svc:
        for c in cases:  # :17-46
            opts = {"sampleRate": c.sampleRate, "fftSize": c.n, "window": "rect", "sides": "one"}
            a, b = spectrum(c.signal, opts), spectrumFx(c.signal, opts)(svc)
            assert a["peak"] == b["peak"]
            assert (a["amplitude"] == b["amplitude"]).all() and (a["phase"] == b["phase"]).all()
            assert (a["frequencies"] == b["frequencies"]).all()
        assert svc.fft(64) is svc.fft(64) and svc.fft(64) is not svc.fft(128)  # :50-70
        w1 = svc.window("hann", 64)
        assert w1 is svc.window("hann", 64) and w1 is not svc.window("hamming", 64) and w1 is not svc.window("hann", 128)
        c = cases[0]  # :98-134
        opts = {"sampleRate": c.sampleRate, "fftSize": c.n, "window": "rect", "sides": "one"}
        frame = c.signal.astype(np.float32)
        res = list(spectrumStream([frame, frame, frame], opts, service=svc))
        exp = spectrum(c.signal, opts)
        assert len(res) == 3
        for r in res:
            assert r["peak"]["index"] == exp["peak"]["index"] and r["peak"]["frequency"] == exp["peak"]["frequency"]
            assert abs(r["peak"]["amplitude"] - exp["peak"]["amplitude"]) < 0.5e-5
        assert list(spectrumStream([], {"sampleRate": 48000, "fftSize": 64}, service=svc)) == []  # :136-146
        for wt in ("rect", "hann", "hamming", "blackman"):  # :149-178
            opts = {"sampleRate": c.sampleRate, "fftSize": c.n, "window": wt, "sides": "one"}
            a, b = spectrum(c.signal, opts), spectrumFx(c.signal, opts)(svc)
            assert a["peak"]["index"] == b["peak"]["index"] and abs(a["peak"]["amplitude"] - b["peak"]["amplitude"]) < 0.5e-10
        # a stream mixing frame lengths keeps order
        mixed = [frame, frame[:512], frame, frame[:512]]
        res = list(spectrumStream(mixed, {"sampleRate": 48000.0}, service=svc, chunk=2))
        assert [len(r["amplitude"]) for r in res] == [513, 257, 513, 257]


def test_apply_window_fftshift_stft_gpu():
    import oracle
    from pragma_dsp_b200 import stft
    from pragma_dsp_b200.xform import applyWindow, createWindow, fftShift
    rng = np.random.default_rng(8)
    x = rng.standard_normal(4096)
    w = createWindow("blackman", 4096)
    assert (applyWindow(x, w) == x * w).all()
    for n in (1, 2, 7, 8, 1025):
        v = rng.standard_normal(n)
        assert (fftShift(v) == oracle.fftShift(v)).all()
    sig = rng.standard_normal(48000).astype(np.float32)
    r = stft(sig, fftSize=4096, hopSize=1024, window="hann", sampleRate=48000.0, outputs=("amplitude", "phase", "peak"))
    frames = (48000 - 4096) // 1024 + 1
    ref = oracle.spectrum_batch(sig, fftSize=4096, frameLen=4096, hop=1024, batch=frames, sampleRate=48000.0, window="hann")
    assert r["amplitude"].shape == (frames, 2049) and np.abs(r["amplitude"] - ref["amplitude"]).max() <= 1e-13
    assert (r["peaks"]["index"] == ref["peaks"]["index"]).all()
