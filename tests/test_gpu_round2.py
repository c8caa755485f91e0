"""Round-2 features on the device (pytest -m gpu): the single-call fast lane, the staged-load kernels, the ingestion
ring's pinned / threaded / polled forms, spectrumStream's latency contract, `out` validation in the host layer."""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def multitone(rng, batch, n, dtype=np.float64):
    t = np.arange(n)
    k = rng.integers(8, max(9, n // 2 - 8), size=(batch, 3)) + rng.uniform(-0.25, 0.25, size=(batch, 3))
    a = np.concatenate([np.ones((batch, 1)), rng.uniform(0.1, 0.5, size=(batch, 2))], axis=1)
    ph = rng.uniform(0, 2 * np.pi, size=(batch, 3))
    x = np.zeros((batch, n))
    for j in range(3):
        x += a[:, j, None] * np.sin(2 * np.pi * k[:, j, None] * t[None, :] / n + ph[:, j, None])
    return x.astype(dtype)


@pytest.fixture(scope="module")
def ctx():
    from pragma_dsp_b200 import _lib
    return _lib.default_context()


@pytest.mark.parametrize("n", [2, 64, 1024, 4096, 16384])
def test_fast_lane_matches_pipeline_and_oracle(ctx, n):
    """One-frame calls go through the fast lane (one launch over the host-mapped buffer, doorbell completion) and give
    the bits of the staging pipeline; both agree with the oracle."""
    from pragma_dsp_b200 import spectrum_batch
    from pragma_dsp_b200.core import ComplexArray, Radix2Fft
    rng = np.random.default_rng(n)
    x = multitone(rng, 3, n) if n >= 64 else rng.standard_normal((3, n))
    fft = Radix2Fft(n)
    f0 = ctx.fast_call_count
    a = fft.forward(x[0])
    c = fft.forwardComplex(ComplexArray(x[1], x[2]))
    b = fft.inverse(c)
    s = spectrum_batch(x[:1], sampleRate=48000.0, fftSize=n, window="hann")
    # the lane takes jobs of up to 256 KB (input + outputs): N = 16384 frames (128 KB in, 128-256 KB out) are beyond it
    fast_jobs = 4 if n <= 8192 else 0
    assert ctx.fast_call_count == f0 + fast_jobs
    try:
        ctx.tune("fast", 0)
        a2 = fft.forward(x[0])
        c2 = fft.forwardComplex(ComplexArray(x[1], x[2]))
        s2 = spectrum_batch(x[:1], sampleRate=48000.0, fftSize=n, window="hann")
        assert ctx.fast_call_count == f0 + fast_jobs
    finally:
        ctx.tune("fast", None)
    assert np.array_equal(a.real, a2.real) and np.array_equal(a.imag, a2.imag)
    assert np.array_equal(c.real, c2.real) and np.array_equal(c.imag, c2.imag)
    for key in ("amplitude", "phase"):
        assert np.array_equal(s[key], s2[key]), key
    assert (s["peaks"] == s2["peaks"]).all()
    rre, rim = oracle.FFT(n).forward(x[0])
    ref = rre.reshape(-1) + 1j * rim.reshape(-1)
    assert np.linalg.norm((a.real + 1j * a.imag) - ref) / np.linalg.norm(ref) <= 1e-12 * max(1, np.log2(n))
    assert np.abs(b.real - x[1]).max() <= 1e-12 and np.abs(b.imag - x[2]).max() <= 1e-12


def test_fast_lane_many_calls_in_a_row(ctx):
    """The doorbell sequence number advances per call: 2,000 back-to-back one-frame calls, each checked."""
    from pragma_dsp_b200.core import Radix2Fft
    n = 1024
    fft = Radix2Fft(n)
    out = fft.forward(np.zeros(n))
    for i in range(2000):
        x = np.zeros(n)
        x[i % n] = 1.0 + i
        fft.forward(x, out)
        assert out.real[0] == 1.0 + i and abs(np.hypot(out.real, out.imag).max() - (1.0 + i)) < 1e-9


@pytest.mark.parametrize("prec,sdtype", [("f64", np.float64), ("f64", np.float32), ("f32", np.float32)])
def test_staged_kernels_are_bit_identical(ctx, prec, sdtype):
    """pdsp_ctx_tune("staged", 1): cp.async.bulk + mbarrier sample loads (MD_STAGED kernels, N = 1024) against the
    direct-load kernels - same arithmetic, same bits; STFT addressing (hop < N) included."""
    from pragma_dsp_b200 import spectrum_batch
    rng = np.random.default_rng(7)
    n = 1024
    x = multitone(rng, 3000, n).astype(sdtype)
    stream = multitone(rng, 1, 40 * 256 + n)[0].astype(sdtype)
    for outs in (("amplitude",), ("amplitude", "peak"), ("peak",), ("amplitude", "phase", "peak")):
        res = []
        for staged in (0, 1):
            ctx.tune("staged", staged)
            try:
                r1 = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann", precision=prec, outputs=outs)
                r2 = spectrum_batch(stream, sampleRate=48000.0, fftSize=n, window="hann", precision=prec, outputs=outs,
                                    frameLen=n, hop=256, batch=41)
            finally:
                ctx.tune("staged", None)
            res.append((r1, r2))
        for a, b in zip(res[0], res[1]):
            for key in ("amplitude", "phase"):
                if a[key] is not None:
                    assert np.array_equal(a[key], b[key]), (outs, key)
            if a["peaks"] is not None:
                assert (a["peaks"] == b["peaks"]).all(), outs


def test_ring_pinned_push_ready_pop_into(ctx):
    from pragma_dsp_b200 import IngestRing, spectrum_batch
    from pragma_dsp_b200._lib import PEAK_F32, check, lib
    L = lib()
    rng = np.random.default_rng(3)
    n, total = 1024, 5000
    x = multitone(rng, total, n, np.float32)
    ref = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann", precision="f32", outputs=("amplitude", "peak"))
    hp = C.c_void_p()
    check(L.pdsp_host_alloc(ctx.h, x.nbytes, C.byref(hp)))
    xp = np.ctypeslib.as_array((C.c_float * x.size).from_address(hp.value)).reshape(x.shape)
    xp[:] = x
    amp = np.empty((total, n // 2 + 1), dtype=np.float32)
    pk = np.zeros(total, dtype=PEAK_F32)
    for pinned in (True, False):
        amp[:] = 0
        with IngestRing(n, sampleRate=48000.0, window="hann", precision="f32", sample_dtype=np.float32,
                        outputs=("amplitude", "peak"), framesPerChunk=1024, depth=3) as ring:
            pushed = got = 0
            while pushed < total:
                blk = xp[pushed:pushed + 700] if pinned else x[pushed:pushed + 700]
                k = ring.push_pinned(blk) if pinned else ring.push(blk)
                pushed += k
                if k < blk.shape[0]:
                    got += ring.pop_into(amp[got:], None, pk[got:], max_frames=1024)
            ring.flush()
            while got < total:
                fin, fly, pend = ring.ready()
                assert fin + fly + pend > 0
                got += ring.pop_into(amp[got:], None, pk[got:], max_frames=min(1500, total - got))
            assert ring.ready() == (0, 0, 0)
        assert np.array_equal(amp, ref["amplitude"]) and (pk == ref["peaks"]).all(), pinned
    check(L.pdsp_host_free(ctx.h, hp))


def test_spectrum_stream_latency_contract(ctx):
    from pragma_dsp_b200.effect import FourierLive, spectrumStream
    rng = np.random.default_rng(41)
    n = 1024
    x = multitone(rng, 40, n, np.float32)
    svc = FourierLive()
    produced, lag = [], []

    def source():
        for i in range(40):
            produced.append(i)
            yield x[i]

    for i, r in enumerate(spectrumStream(source(), {"sampleRate": 8000.0}, service=svc, chunk=4096)):
        lag.append(len(produced) - i)
    assert len(lag) == 40 and max(lag) <= 8, lag     # never `chunk` (4096) frames behind: bounded by what is in flight
    seen = []

    def feedback():
        for i in range(6):
            assert len(seen) == i
            yield x[i]

    for r in spectrumStream(feedback(), {"sampleRate": 8000.0}, service=svc, lockstep=True):
        seen.append(r["peak"]["index"])
    assert seen == [oracle.spectrum(x[i], sampleRate=8000.0)["peak"]["index"] for i in range(6)]


def test_out_validation_before_native_code(ctx):
    from pragma_dsp_b200.core import ComplexArray, Radix2Fft
    fft = Radix2Fft(64)
    x = np.arange(64.0)
    with pytest.raises(ValueError, match="length 64"):
        fft.forward(x, ComplexArray(np.zeros(32), np.zeros(64)))
    with pytest.raises(ValueError, match="contiguous"):
        fft.forward(x, ComplexArray(np.zeros(128)[::2], np.zeros(64)))
    with pytest.raises(TypeError, match="float64"):
        fft.forward(x, ComplexArray(np.zeros(64, dtype=np.float32), np.zeros(64)))
    out = ComplexArray(np.zeros(64), np.zeros(64))
    assert fft.forward(x, out) is out and out.real[0] == x.sum()      # `out` identity (src/core/fft.ts:106,150)
