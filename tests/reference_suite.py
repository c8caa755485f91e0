"""The reference's own test suite, restated once and run against any implementation.

Each ``check_*`` function replays one ``describe``/``it`` block of the reference's vitest files
(cited per function) with the reference's tolerances, against an ``impl`` adaptor:

    impl.forward(n, x)            -> (re, im)      new FFT(n).forward(x)
    impl.inverse(n, re, im)       -> (re, im)      new FFT(n).inverse({real, imag})
    impl.magnitude(re, im)        -> array         magnitude()
    impl.phase(re, im)            -> array         phase()
    impl.createWindow(type, n)    -> array
    impl.binFrequencies(n, fs, sides)
    impl.spectrum(x, sampleRate=, fftSize=, window=, sides=) -> dict(frequencies, amplitude, phase, peak{...})

tests/test_oracle_golden.py binds it to the CPU oracle (pins the oracle, runs without a GPU);
tests/test_gpu_golden.py binds it to the CUDA path through the C-ABI (``-m gpu``).
"""
from __future__ import annotations

import math

import numpy as np

from conftest import FIXTURE_CASES, FIXTURE_WINDOWS, REALLIFE_WINDOWS, reallife


def expect_close_array(actual, expected, tol):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, (actual.shape, expected.shape)
    diff = np.abs(actual - expected)
    i = int(np.argmax(diff)) if diff.size else 0
    assert not np.isnan(diff).any(), "NaN in comparison"
    assert diff.size == 0 or diff[i] <= tol, f"mismatch at {i}: actual={actual[i]!r} expected={expected[i]!r} diff={diff[i]} tol={tol}"


def to_be_close_to(actual, expected, digits):
    """vitest toBeCloseTo(expected, digits): |a-e| < 10^-digits / 2."""
    assert abs(actual - expected) < 10.0 ** (-digits) / 2, (actual, expected, digits)


def wrapped_phase_diff(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b))
    return np.minimum(d, np.abs(d - 2 * math.pi))


# ---------------------------------------------------------------- test/fft.test.ts:18-43
def check_fft_fixtures(impl, tol=1e-6):
    n_checked = 0
    for n in (8, 16, 32):
        for c in [c for c in FIXTURE_CASES if c.n == n and c.kind == "random_normal"]:
            re, im = impl.forward(c.n, c.input)
            expect_close_array(re, c.fftRe, tol)
            expect_close_array(im, c.fftIm, tol)
            rr, ri = impl.inverse(c.n, re, im)
            expect_close_array(rr, c.input, tol)
            expect_close_array(ri, np.zeros(c.n), tol)
            n_checked += 1
    assert n_checked == 15


# ---------------------------------------------------------------- test/spectrum.test.ts:15-34
def check_spectrum_fixture(impl):
    (c,) = [c for c in FIXTURE_CASES if c.kind == "sine_bin_centered"]
    r = impl.spectrum(c.input, sampleRate=c.sampleRate, fftSize=c.n, window="rect", sides="one")
    assert r["peak"]["index"] == c.meta["binCenteredK"]
    assert abs(r["peak"]["frequency"] - c.meta["expectedPeakHz"]) <= 1e-6
    assert abs(r["peak"]["amplitude"] - c.meta["amplitude"]) <= 1e-3


# ---------------------------------------------------------------- test/window.test.ts:21-25
def check_window_fixtures(impl, tol=1e-8):
    assert len(FIXTURE_WINDOWS) == 28
    for w in FIXTURE_WINDOWS:
        expect_close_array(impl.createWindow(w.type, w.n), w["values"], tol)
    # test/reallife/references/windows_dsp.json (loaded by nobody in the reference; same generator)
    for w in REALLIFE_WINDOWS:
        expect_close_array(impl.createWindow(w.type, w.n), w["values"], tol)


# ---------------------------------------------------------------- bench/run.ts:20-35
def check_bench_checksums(impl, checksum_fn):
    want = {2048: "-1.340946", 4096: "-5.113685"}
    for c in [c for c in FIXTURE_CASES if c.kind == "benchmark_random_normal"]:
        re, im = impl.forward(c.n, c.input)
        assert f"{checksum_fn(re, im):.6f}" == want[c.n]


# ---------------------------------------------------------------- test/reallife/signals.test.ts
def check_signals_pure_sine(impl, tol=1e-10):
    """:12-66 - FFT, magnitude, masked wrap-aware phase, round trip for the 23 pure-sine cases."""
    cases = reallife("pure_sine")
    assert len(cases) == 23
    for c in cases:
        re, im = impl.forward(c.n, c.signal)
        expect_close_array(re, c.fftRe, tol)
        expect_close_array(im, c.fftIm, tol)
        expect_close_array(impl.magnitude(re, im), c.magnitude, tol)
        ph = impl.phase(re, im)
        mask = c.magnitude > 1e-6
        assert (wrapped_phase_diff(ph[mask], c.phase[mask]) < tol).all(), c.name
        rr, ri = impl.inverse(c.n, re, im)
        expect_close_array(rr, c.signal, tol)
        assert np.max(np.abs(ri)) < tol


def check_signals_multi_tone(impl, tol=1e-10):
    """:68-98"""
    cases = reallife("multi_tone")
    assert len(cases) == 2
    for c in cases:
        re, im = impl.forward(c.n, c.signal)
        expect_close_array(re, c.fftRe, tol)
        expect_close_array(im, c.fftIm, tol)
        mag = impl.magnitude(re, im)
        for b, a in zip(c.params["bin_indices"], c.params["amplitudes"]):
            to_be_close_to(mag[b], c.n * a / 2, 5)


def check_signals_chirp(impl, tol=1e-10):
    """:100-120"""
    for c in reallife("chirp"):
        re, im = impl.forward(c.n, c.signal)
        expect_close_array(re, c.fftRe, tol)
        expect_close_array(im, c.fftIm, tol)
        rr, _ = impl.inverse(c.n, re, im)
        expect_close_array(rr, c.signal, tol)


def check_signals_special(impl):
    """:122-196 - impulse flat, DC only bin 0, Nyquist only bin N/2, zeros exactly 0."""
    imp = [c for c in reallife("special", "impulse") if c.params["position"] == 0][0]
    mag = impl.magnitude(*impl.forward(imp.n, imp.signal))
    for v in mag:
        to_be_close_to(v, imp.params["amplitude"], 10)
    (dc,) = reallife("special", "dc")
    mag = impl.magnitude(*impl.forward(dc.n, dc.signal))
    to_be_close_to(mag[0], dc.n * dc.params["level"], 10)
    assert (mag[1:] < 1e-10).all()
    (ny,) = reallife("special", "nyquist")
    mag = impl.magnitude(*impl.forward(ny.n, ny.signal))
    to_be_close_to(mag[ny.n // 2], ny.n * ny.params["amplitude"], 10)
    others = np.delete(mag, ny.n // 2)
    assert (others < 1e-10).all()
    (z,) = reallife("special", "zeros")
    re, im = impl.forward(z.n, z.signal)
    assert (np.asarray(re) == 0).all() and (np.asarray(im) == 0).all()


# ---------------------------------------------------------------- test/reallife/scaling.test.ts
def check_scaling(impl):
    centered = reallife("pure_sine", "pure_sine_bin_centered")
    assert len(centered) == 15
    for c in centered:
        r = impl.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, window="rect", sides="one")  # :14-33
        assert r["peak"]["index"] == c.params["bin_index"], c.name
        to_be_close_to(r["peak"]["amplitude"], c.params["amplitude"], 2)
        to_be_close_to(r["peak"]["frequency"], c.params["frequency_hz"], 6)  # :119-136
        r2 = impl.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, window="rect", sides="two")  # :80-100
        k = c.params["bin_index"]
        to_be_close_to(r2["amplitude"][k], c.params["amplitude"] / 2, 2)
        to_be_close_to(r2["amplitude"][c.n - k], c.params["amplitude"] / 2, 2)
    (dc,) = reallife("special", "dc")
    r = impl.spectrum(dc.signal, sampleRate=dc.sampleRate, fftSize=dc.n, window="rect", sides="one")
    to_be_close_to(r["amplitude"][0], dc.params["level"], 6)  # :35-49 DC not doubled
    assert r["peak"]["index"] == 0  # :185-201 pure DC -> index 0
    (ny,) = reallife("special", "nyquist")
    r = impl.spectrum(ny.signal, sampleRate=ny.sampleRate, fftSize=ny.n, window="rect", sides="one")
    to_be_close_to(r["amplitude"][ny.n // 2], ny.params["amplitude"], 6)  # :51-70 Nyquist not doubled
    c = reallife("pure_sine")[0]
    r = impl.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, window="rect", sides="two")  # :102-116
    assert len(r["amplitude"]) == c.n and len(r["frequencies"]) == c.n and len(r["phase"]) == c.n
    r = impl.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, window="rect", sides="one")  # :138-165
    assert r["frequencies"][0] == 0
    bw = c.sampleRate / c.n
    for i, f in enumerate(r["frequencies"]):
        to_be_close_to(f, i * bw, 10)
    to_be_close_to(r["frequencies"][-1], c.sampleRate / 2, 10)
    (dps,) = reallife("special", "dc_plus_sine")  # :168-183 peak ignores DC
    r = impl.spectrum(dps.signal, sampleRate=dps.sampleRate, fftSize=dps.n, window="rect", sides="one")
    assert r["peak"]["index"] == dps.params["sine_bin"]


# ---------------------------------------------------------------- test/reallife/phase.test.ts
def check_phase(impl):
    sine8 = [c for c in reallife("pure_sine", "pure_sine_bin_centered") if c.params["bin_index"] == 8][0]
    (cos8,) = reallife("cosine")
    ps = impl.phase(*impl.forward(sine8.n, sine8.signal))
    pc = impl.phase(*impl.forward(cos8.n, cos8.signal))
    d = pc[8] - ps[8]  # :7-43 cosine leads sine by pi/2
    while d > math.pi:
        d -= 2 * math.pi
    while d < -math.pi:
        d += 2 * math.pi
    assert abs(d - math.pi / 2) < 1e-6
    phase_cases = reallife("pure_sine", "pure_sine_phase")
    assert len(phase_cases) == 5
    for c in phase_cases:
        b = c.params["bin_index"]
        ph = impl.phase(*impl.forward(c.n, c.signal))  # :45-73
        assert wrapped_phase_diff(ph[b], c.phase[b]) < 1e-10
        r = impl.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, window="rect", sides="one")  # :75-100
        assert r["peak"]["index"] == b
        assert wrapped_phase_diff(r["peak"]["phase"], c.phase[b]) < 1e-10
    c = reallife("pure_sine")[0]  # :102-134
    assert len(impl.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, sides="one")["phase"]) == c.n // 2 + 1
    assert len(impl.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n, sides="two")["phase"]) == c.n
    ph = impl.phase(*impl.forward(64, np.full(64, 1.0)))  # :136-161
    to_be_close_to(ph[0], 0.0, 10)
    ph = impl.phase(*impl.forward(64, np.full(64, -1.0)))
    to_be_close_to(abs(ph[0]), math.pi, 10)


# ---------------------------------------------------------------- test/reallife/edge_cases.test.ts
def check_edge_cases(impl):
    r = impl.spectrum(np.zeros(64), sampleRate=48000, fftSize=64, window="rect", sides="one")  # :22-38
    assert (np.asarray(r["amplitude"]) == 0).all() and r["peak"]["amplitude"] == 0
    assert r["peak"]["index"] == 0
    imp = [c for c in reallife("special", "impulse") if c.params["position"] > 0][0]  # :110-126
    for v in impl.magnitude(*impl.forward(imp.n, imp.signal)):
        to_be_close_to(v, imp.params["amplitude"], 10)
    (tiny,) = reallife("special", "tiny")  # :129-148
    re, im = impl.forward(tiny.n, tiny.signal)
    assert np.isfinite(re).all() and np.isfinite(im).all()
    assert np.max(np.abs(re - tiny.fftRe)) < 1e-20
    (large,) = reallife("special", "large")  # :150-177
    re, im = impl.forward(large.n, large.signal)
    assert np.isfinite(re).all() and np.isfinite(im).all()
    big = np.abs(large.fftRe) > 1
    assert (np.abs(re[big] - large.fftRe[big]) / np.abs(large.fftRe[big]) < 1e-9).all()
    assert (np.abs(re[~big] - large.fftRe[~big]) < 1e-6).all()
    r = impl.spectrum(np.array([1.0, 2, 3, 4]), sampleRate=48000, fftSize=16, window="rect", sides="one")  # :179-197
    assert len(r["amplitude"]) == 9 and math.isfinite(r["peak"]["amplitude"]) and math.isfinite(r["peak"]["frequency"])
    r = impl.spectrum(np.array([1.0, 1, 1, 1]), sampleRate=48000, fftSize=16, window="rect", sides="one")  # :199-213
    to_be_close_to(r["amplitude"][0], 4 / 16, 6)
    for c in reallife("special"):  # :216-235
        re, im = impl.forward(c.n, c.signal)
        rr, ri = impl.inverse(c.n, re, im)
        assert np.max(np.abs(rr - c.signal)) < 1e-9 and np.max(np.abs(ri)) < 1e-9


# ---------------------------------------------------------------- README.md:43-50 (BASELINE config C1)
def check_readme_core_example(impl):
    x = np.sin(np.arange(1024, dtype=np.float64))
    re, im = impl.forward(1024, x)
    ref = np.fft.fft(x)
    rel = np.linalg.norm((re + 1j * im) - ref) / np.linalg.norm(ref)
    assert rel <= 1e-12 * 10, rel  # north_star: 1e-12*log2(N)
    r = impl.spectrum(x, fftSize=1024)
    assert r["peak"]["index"] == 163


ALL_CHECKS = [
    check_fft_fixtures, check_spectrum_fixture, check_window_fixtures, check_signals_pure_sine,
    check_signals_multi_tone, check_signals_chirp, check_signals_special, check_scaling, check_phase,
    check_edge_cases, check_readme_core_example,
]
