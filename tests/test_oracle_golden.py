"""Pins the CPU oracle (oracle/pragma_oracle.c) against every golden vector the reference's
own tests hold for the hot path.  Runs without a GPU."""
import numpy as np
import pytest

import oracle
import reference_suite as suite


class OracleImpl:
    @staticmethod
    def forward(n, x):
        return oracle.FFT(n).forward(np.asarray(x, dtype=np.float64))

    @staticmethod
    def inverse(n, re, im):
        return oracle.FFT(n).inverse(re, im)

    magnitude = staticmethod(oracle.magnitude)
    phase = staticmethod(oracle.phase)
    createWindow = staticmethod(oracle.createWindow)
    binFrequencies = staticmethod(oracle.binFrequencies)

    @staticmethod
    def spectrum(x, sampleRate=1.0, fftSize=None, window="rect", sides="one"):
        return oracle.spectrum(x, sampleRate=sampleRate, fftSize=fftSize, window=window, sides=sides)


@pytest.mark.parametrize("check", suite.ALL_CHECKS, ids=lambda f: f.__name__)
def test_oracle_replays_reference_suite(check):
    check(OracleImpl)


def test_oracle_bench_run_checksums():
    """bench/run.ts:20-35 guardrail output on the regenerated bench_rand_n{2048,4096} inputs."""
    suite.check_bench_checksums(OracleImpl, oracle.bench_checksum)


def test_oracle_pow2_helpers():
    """src/core/fft.ts:16-23"""
    assert [oracle.nextPowerOfTwo(n) for n in (-3, 0, 1, 2, 3, 4, 5, 1023, 1024, 1025)] == [1, 1, 1, 2, 4, 4, 8, 1024, 1024, 2048]
    assert [oracle.isPowerOfTwo(n) for n in (-4, 0, 1, 2, 3, 4, 6, 1024)] == [False, False, True, True, False, True, False, True]
    with pytest.raises(ValueError, match="FFT size must be power of two, got 12"):
        oracle.FFT(12)


def test_oracle_window_semantics():
    """src/xform/fourier.ts:14-52 and SURVEY Appendix A.3"""
    assert oracle.createWindow("hann", 1).tolist() == [1.0]
    h = oracle.createWindow("hann", 1024)
    assert h[0] == 0.0 and abs(h[-1]) < 1e-15
    assert oracle.createWindow("blackman", 1024)[0] == -1.3877787807814457e-17
    with pytest.raises(ValueError):
        oracle.createWindow("hann", 0)
    with pytest.raises(ValueError):
        oracle.createWindow("kaiser", 8)


def test_oracle_two_sided_mirror_peaks_documented():
    """with sides="two" the reference lands on the mirror bin N-k in exactly 3 of the
    35 golden cases on 1-ulp noise.  Pin that the oracle reproduces this (it is what makes the
    two-sided index parity rule `index in {k, N-k}` necessary for a different FFT algorithm)."""
    from conftest import REALLIFE_CASES
    mirrored = []
    for c in REALLIFE_CASES:
        one = oracle.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, sides="one")["peak"]["index"]
        two = oracle.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, sides="two")["peak"]["index"]
        assert two in (one, (c.n - one) % c.n), c.name
        if two != one:
            mirrored.append(c.name)
    assert sorted(mirrored) == ["chirp_100hz_to_2000hz", "sine_2500hz", "sine_bin8_phase45deg"]


def test_oracle_batch_stft_addressing_and_threads():
    rng = np.random.default_rng(1337)
    sig = rng.standard_normal(4096 + 7 * 1024).astype(np.float32)
    a = oracle.spectrum_batch(sig, fftSize=4096, frameLen=4096, hop=1024, batch=8, sampleRate=48000.0, window="hann")
    b = oracle.spectrum_batch(sig, fftSize=4096, frameLen=4096, hop=1024, batch=8, sampleRate=48000.0, window="hann", threads=4)
    assert (a["amplitude"] == b["amplitude"]).all() and (a["peaks"] == b["peaks"]).all()
    for f in (0, 3, 7):
        one = oracle.spectrum(sig[f * 1024:f * 1024 + 4096], sampleRate=48000.0, fftSize=4096, window="hann")
        assert (one["amplitude"] == a["amplitude"][f]).all() and one["peak"]["index"] == a["peaks"][f]["index"]
