"""Runs the REAL kernel source (pragma_dsp_b200/csrc/fft_kernels.cuh) under the host SIMT
emulator (tests/simt_emu) and checks it against the oracle.  No GPU needed.

This is a development/verification harness for index math, synchronisation and the fused
epilogue; the `-m gpu` tests are the parity tests proper (same checks, through the C-ABI on a B200).
"""
import numpy as np
import pytest

import oracle
import simt_emu as E

# north_star tolerances
def tol(n, dtype):
    return 1e-12 * max(1.0, np.log2(n)) if dtype == np.float64 else 1e-5


def rel_l2(a, b):
    d = np.linalg.norm(np.asarray(a, dtype=np.complex128) - b)
    s = np.linalg.norm(b)
    return d / s if s > 0 else d


def multitone(rng, batch, n):
    t = np.arange(n)
    out = np.empty((batch, n))
    for f in range(batch):
        k = rng.integers(1, max(2, n // 2 - 1), size=3) + rng.uniform(-0.25, 0.25, size=3)
        a = np.array([1.0, rng.uniform(0.1, 0.5), rng.uniform(0.1, 0.5)])
        ph = rng.uniform(0, 2 * np.pi, size=3)
        out[f] = (a[:, None] * np.sin(2 * np.pi * k[:, None] * t[None, :] / n + ph[:, None])).sum(0)
    return out


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_r2c_forward_matches_oracle(n, dtype):
    rng = np.random.default_rng(n)
    batch = 5
    x = rng.standard_normal((batch, n))
    if dtype == np.float32:
        x = x.astype(np.float32)
    o = E.r2c(x.reshape(-1), n, dtype=dtype, batch=batch, want=("complex",), nblocks=2)
    ore, oim = oracle.FFT(n).forward(x.astype(np.float64))
    for f in range(batch):
        assert rel_l2(o["re"][f] + 1j * o["im"][f], ore[f] + 1j * oim[f]) <= tol(n, dtype)


@pytest.mark.parametrize("n", [1, 2, 4, 8, 64, 512, 1024, 2048])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("inverse", [False, True])
def test_c2c_matches_oracle(n, dtype, inverse):
    rng = np.random.default_rng(100 + n)
    batch = 3
    re = rng.standard_normal((batch, n))
    im = rng.standard_normal((batch, n))
    if dtype == np.float32:
        re, im = re.astype(np.float32), im.astype(np.float32)
    ore, oim = E.c2c(re, im, n, dtype=dtype, inverse=inverse, nblocks=2)
    plan = oracle.FFT(n)
    rre, rim = (plan.inverse if inverse else plan.forwardComplex)(re.astype(np.float64), im.astype(np.float64))
    for f in range(batch):
        assert rel_l2(ore[f] + 1j * oim[f], rre[f] + 1j * rim[f]) <= tol(n, dtype)


@pytest.mark.parametrize("n,window,sides", [(1024, "hann", "one"), (1024, "rect", "two"), (256, "blackman", "one"),
                                            (64, "hamming", "two"), (4096, "hann", "one"), (16, "hann", "one"),
                                            (8, "rect", "one"), (2, "rect", "one"), (4, "hann", "two")])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_spectrum_epilogue_matches_oracle(n, window, sides, dtype):
    rng = np.random.default_rng(7 * n)
    batch = 6
    x = multitone(rng, batch, n) if n >= 32 else rng.standard_normal((batch, n))
    xs = x.astype(np.float32) if dtype == np.float32 else x
    w = oracle.createWindow(window, n)
    o = E.r2c(xs.reshape(-1), n, dtype=dtype, batch=batch, window=None if window == "rect" else w, sides=sides,
              sample_rate=48000.0, want=("amp", "phase", "peak"), nblocks=2)
    ref = oracle.spectrum_batch(xs, fftSize=n, sampleRate=48000.0, window=window, sides=sides)
    atol = 1e-12 if dtype == np.float64 else 2e-6
    for f in range(batch):
        scale = max(1.0, np.abs(ref["amplitude"][f]).max())
        assert np.abs(o["amp"][f] - ref["amplitude"][f]).max() <= atol * scale * 10
        mask = ref["amplitude"][f] > (1e-6 if dtype == np.float64 else 1e-3) * scale
        d = np.abs(o["phase"][f] - ref["phase"][f])
        d = np.minimum(d, np.abs(d - 2 * np.pi))
        assert d[mask].max(initial=0) <= (1e-9 if dtype == np.float64 else 2e-3)
        pk, rk = o["peaks"][f], ref["peaks"][f]
        if sides == "one":
            assert pk["index"] == rk["index"]
        else:  # mirror bins tie exactly in this implementation; the reference may pick N-k on 1-ulp noise
            assert pk["index"] in (rk["index"], (n - rk["index"]) % n)
        assert abs(pk["amplitude"] - rk["amplitude"]) <= atol * scale * 10
        assert pk["frequency"] == pytest.approx(pk["index"] * (48000.0 / n), rel=1e-6 if dtype == np.float32 else 1e-15)
        assert pk["amplitude"] == o["amp"][f][pk["index"]]
        assert pk["phase"] == pytest.approx(o["phase"][f][pk["index"]], abs=1e-6)


def test_frame_addressing_zero_pad_truncate_hop():
    """buildFrame semantics (src/public/spectrum.ts:36-43) + STFT hop addressing, f32 samples into an f64 plan."""
    rng = np.random.default_rng(3)
    sig = rng.standard_normal(5000).astype(np.float32)
    for n, frame_len, hop, batch in [(256, 100, 37, 9), (256, 256, 64, 12), (128, 300, 301, 7), (64, 0, 5, 3), (512, 511, 1, 5)]:
        o = E.r2c(sig, n, dtype=np.float64, frame_len=frame_len, hop=hop, batch=batch, want=("amp", "peak"),
                  window=oracle.createWindow("hann", n), sample_rate=8000.0, nblocks=2)
        ref = oracle.spectrum_batch(sig, fftSize=n, frameLen=frame_len, hop=hop, batch=batch, sampleRate=8000.0, window="hann")
        assert np.abs(o["amp"] - ref["amplitude"]).max() <= 1e-13
        assert (o["peaks"]["index"] == ref["peaks"]["index"]).all()


def test_exactness_rules():
    """zeros -> exactly 0 and peak index 0; constant + rect -> non-DC bins exactly 0 and peak index 0
    (test/reallife/scaling.test.ts:185-201, edge_cases.test.ts:22-38)."""
    for n in (2, 4, 8, 64, 1024):
        for dtype in (np.float64, np.float32):
            o = E.r2c(np.zeros(n, dtype=dtype), n, dtype=dtype, sample_rate=48000.0)
            assert (o["re"] == 0).all() and (o["im"] == 0).all() and (o["amp"] == 0).all()
            assert o["peaks"][0]["index"] == 0 and o["peaks"][0]["amplitude"] == 0 and o["peaks"][0]["phase"] == 0
            o = E.r2c(np.full(n, 0.75, dtype=dtype), n, dtype=dtype, sample_rate=48000.0)
            assert (o["amp"][0][1:] == 0).all() and o["amp"][0][0] == 0.75
            assert o["peaks"][0]["index"] == 0 and o["peaks"][0]["amplitude"] == 0.75
            o = E.r2c(np.full(n, -0.5, dtype=dtype), n, dtype=dtype)
            assert o["peaks"][0]["index"] == 0 and abs(o["peaks"][0]["phase"]) == pytest.approx(np.pi)


def test_peak_tie_breaks_to_lowest_index():
    """Two exactly equal bins: strict '>' keeps the first (src/public/spectrum.ts:81-84)."""
    n = 64
    t = np.arange(n)
    x = np.cos(2 * np.pi * 5 * t / n) + np.cos(2 * np.pi * 20 * t / n)  # equal amplitude at bins 5 and 20
    o = E.r2c(x, n, want=("amp", "peak"))
    ref = oracle.spectrum(x, fftSize=n)
    assert ref["peak"]["index"] in (5, 20)
    if o["amp"][0][5] == o["amp"][0][20]:
        assert o["peaks"][0]["index"] == 5
    # Nyquist-only signal: index N/2, halved scale (scaling.test.ts:51-70)
    x = np.where(t % 2 == 0, 1.0, -1.0)
    o = E.r2c(x, n, want=("amp", "peak"))
    assert o["peaks"][0]["index"] == n // 2 and o["peaks"][0]["amplitude"] == 1.0


def test_extreme_magnitudes_use_hypot_path():
    n = 64
    t = np.arange(n)
    for scale in (1e-170, 1e170):
        x = scale * np.sin(2 * np.pi * 3 * t / n)
        o = E.r2c(x, n, raw=True, want=("amp", "peak"))
        ref = oracle.magnitude(*oracle.FFT(n).forward(x))
        assert np.isfinite(o["amp"]).all()
        assert abs(o["amp"][0][3] - ref[3]) <= 1e-12 * ref[3]
        assert o["peaks"][0]["index"] == 3


def test_nan_never_wins_peak():
    n = 32
    x = np.sin(2 * np.pi * 4 * np.arange(n) / n)
    x[7] = np.nan  # every bin becomes NaN: no v > 0 -> index 0 (findPeak comparisons are all false)
    o = E.r2c(x, n, want=("amp", "peak"))
    ref = oracle.spectrum(x, fftSize=n)
    assert o["peaks"][0]["index"] == ref["peak"]["index"] == 0


def test_golden_reallife_through_emulated_kernel():
    """The 35 NumPy golden cases (N=1024) through the real kernel source, reference tolerances."""
    from conftest import REALLIFE_CASES
    sig = np.stack([c.signal for c in REALLIFE_CASES])
    o = E.r2c(sig.reshape(-1), 1024, batch=len(REALLIFE_CASES), sample_rate=48000.0, nblocks=3)
    for i, c in enumerate(REALLIFE_CASES):
        scale = c.params["amplitude"] if c.kind == "large" else 1.0  # edge_cases.test.ts:150-177 is relative
        assert np.abs(o["re"][i] - c.fftRe).max() <= 1e-10 * scale, c.name
        assert np.abs(o["im"][i] - c.fftIm).max() <= 1e-10 * scale, c.name
        ref = oracle.spectrum(c.signal, sampleRate=c.sampleRate, fftSize=c.n)
        assert o["peaks"][i]["index"] == ref["peak"]["index"], c.name


@pytest.mark.parametrize("n", [64, 256, 1024, 4096])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("want", [("amp",), ("amp", "peak"), ("peak",), ("complex",), ("amp", "phase", "peak")])
def test_specialised_kernels_equal_generic(n, dtype, want):
    """The compile-time specialised kernels (MD_* modes) must produce the generic kernel's bits."""
    rng = np.random.default_rng(n + len(want))
    batch = 7  # not a multiple of the CTA's frame slots: exercises the tail slot
    x = multitone(rng, batch, n).astype(dtype)
    w = oracle.createWindow("hann", n)
    kw = dict(dtype=dtype, batch=batch, window=w, sample_rate=48000.0, want=want, nblocks=2)
    a = E.r2c(x.reshape(-1), n, **kw)
    b = E.r2c(x.reshape(-1), n, specialised=True, **kw)
    for key in a:
        if key == "peaks":
            assert (a[key] == b[key]).all()
        else:
            assert np.array_equal(a[key], b[key]), key
    # rect window path of the specialised load
    kw["window"] = None
    a = E.r2c(x.reshape(-1), n, **kw)
    b = E.r2c(x.reshape(-1), n, specialised=True, **kw)
    for key in a:
        assert (a[key] == b[key]).all()


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 8, 11])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tuning_variants_match_oracle(variant, dtype):
    """Alternative thread/radix mappings of N=1024 (fft_config.h PDSP_VARIANT): radix-16 and radix-32
    passes, 16-thread frames, two-warp frames with named barriers - same answers."""
    n, batch = 1024, 11
    rng = np.random.default_rng(variant)
    x = multitone(rng, batch, n).astype(dtype)
    w = oracle.createWindow("hann", n)
    ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window="hann")
    o = E.r2c(x.reshape(-1), n, dtype=dtype, batch=batch, window=w, sample_rate=48000.0, want=("amp", "peak"),
              nblocks=2, specialised=True, variant=variant)
    atol = 1e-13 if dtype == np.float64 else 3e-6
    assert np.abs(o["amp"] - ref["amplitude"]).max() <= atol
    assert (o["peaks"]["index"] == ref["peaks"]["index"]).all()
    o2 = E.r2c(x.reshape(-1), n, dtype=dtype, batch=batch, window=w, sample_rate=48000.0, want=("peak",),
               nblocks=1, specialised=True, variant=variant)
    assert (o2["peaks"] == o["peaks"]).all()


@pytest.mark.parametrize("n,dtype,sdtype", [(1024, np.float64, np.float64), (1024, np.float64, np.float32),
                                            (1024, np.float32, np.float32), (256, np.float64, np.float64),
                                            (64, np.float32, np.float32)])
@pytest.mark.parametrize("want", [("amp",), ("amp", "peak"), ("peak",), ("complex",), ("amp", "phase", "peak")])
def test_staged_loads_match_direct_loads(n, dtype, sdtype, want):
    """MD_STAGED kernels fetch the next frame of a slot with a bulk asynchronous copy (cp.async.bulk + mbarrier) into
    the slot's exchange buffer while the current frame is transformed.  Same arithmetic as the direct-load kernels:
    results must be bit-identical, for batches that leave tail slots, several iterations per CTA and overlapping
    (hop < N) frames."""
    rng = np.random.default_rng(n)
    for batch, hop, nblocks in ((11, n, 2), (3, n, 1), (9, n // 4, 3)):
        x = multitone(rng, 1, (batch - 1) * hop + n)[0].astype(sdtype)
        w = oracle.createWindow("hann", n)
        kw = dict(dtype=dtype, batch=batch, hop=hop, frame_len=n, window=w, sample_rate=48000.0, want=want,
                  nblocks=nblocks, specialised=True)
        a = E.r2c(x, n, **kw)
        b = E.r2c(x, n, staged=True, **kw)
        for key in a:
            assert (a[key] == b[key]).all(), (key, batch, hop)
    if "amp" in want:
        frames = np.stack([x[f * hop:f * hop + n] for f in range(batch)])
        ref = oracle.spectrum_batch(frames, fftSize=n, sampleRate=48000.0, window="hann")
        assert np.abs(b["amp"] - ref["amplitude"]).max() <= (1e-13 if dtype == np.float64 else 3e-6)


@pytest.mark.parametrize("n", [64, 1024])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("want", [("amp",), ("amp", "phase", "peak")])
def test_two_sided_specialised_equals_generic(n, dtype, want):
    """MD_TWO: the two-sided rows (mirror bins, optional fused fftShift) of the specialised kernels carry the generic
    kernel's bits."""
    rng = np.random.default_rng(n + 7)
    batch = 5
    x = multitone(rng, batch, n).astype(dtype)
    w = oracle.createWindow("hamming", n)
    kw = dict(dtype=dtype, batch=batch, window=w, sample_rate=8000.0, want=want, nblocks=2, sides="two")
    a = E.r2c(x.reshape(-1), n, **kw)
    b = E.r2c(x.reshape(-1), n, specialised=True, **kw)
    for key in a:
        assert (a[key] == b[key]).all(), key


@pytest.mark.parametrize("n", [64, 1024])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("want", [("amp", "peak"), ("amp", "phase", "peak")])
def test_zero_padded_specialised_equals_generic(n, dtype, want):
    """MD_PAD: frames shorter than N (odd and even lengths, a single sample) through the specialised kernels'
    predicated loads carry the generic kernel's bits."""
    rng = np.random.default_rng(n + 11)
    batch = 5
    w = oracle.createWindow("hann", n)
    for flen in (n - 4, n // 2 + 2, 2):  # even lengths: odd ones stay on the generic kernel
        hop = flen
        buf = rng.standard_normal((batch - 1) * hop + flen).astype(dtype)
        kw = dict(dtype=dtype, batch=batch, window=w, sample_rate=8000.0, want=want, nblocks=2, frame_len=flen, hop=hop)
        a = E.r2c(buf, n, **kw)
        b = E.r2c(buf, n, specialised=True, **kw)
        for key in a:
            assert (a[key] == b[key]).all(), (flen, key)
