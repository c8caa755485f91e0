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


def test_c5_full_size_peak_indices_are_exact(ctx):
    """BASELINE config C5 at its full size on one GPU: 2^20 frames x N=1024 fp64, Hann window, peak only.  Every frame is
    a pure tone on an integer bin (known answer: peak index == that bin, frequency == bin * fs / N); a strided subset is also compared with
    the oracle record for record."""
    import torch
    from pragma_dsp_b200._lib import F64, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib
    L = lib()
    n, frames, fs = 1024, 1 << 20, 48000.0
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    bins = torch.randint(1, n // 2, (frames,), generator=g, device=dev)
    amp = torch.rand(frames, generator=g, device=dev, dtype=torch.float64) + 0.5
    ph = torch.rand(frames, generator=g, device=dev, dtype=torch.float64) * 6.283185307179586
    t = torch.arange(n, device=dev, dtype=torch.float64)
    x = torch.empty((frames, n), dtype=torch.float64, device=dev)
    for f0 in range(0, frames, 1 << 17):  # built in slices: the phase matrix of all frames at once would be 8 GB
        sl = slice(f0, f0 + (1 << 17))
        x[sl] = amp[sl, None] * torch.cos(2 * torch.pi * bins[sl, None].double() * t[None, :] / n + ph[sl, None])
    peaks = torch.zeros(frames * PEAK_F64.itemsize, dtype=torch.uint8, device=dev)
    d = SpectrumDesc(sample_rate=fs, frame_len=n, hop=n, batch=frames, window=WINDOWS["hann"], sides=SIDES["one"],
                     sample_dtype=F64, raw_magnitude=0, fft_shift=0)  # outputs = the non-null pointers: peaks only
    plan = ctx.plan(n, F64)
    st = torch.cuda.Stream(device=dev)
    st.wait_stream(torch.cuda.current_stream())  # the library's launch is ordered behind the generation and the zero fill
    check(L.pdsp_spectrum_dev(plan, C.byref(d), C.c_void_p(x.data_ptr()), None, None, C.c_void_p(peaks.data_ptr()),
                              C.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    rec = peaks.cpu().numpy().view(PEAK_F64).reshape(-1)
    want = bins.cpu().numpy()
    assert (rec["index"] == want).all()
    assert np.abs(rec["frequency"] - want * fs / n).max() == 0.0
    idx = np.arange(0, frames, 4099)
    ref = oracle.spectrum_batch(x[torch.as_tensor(idx, device=dev)].cpu().numpy(), fftSize=n, sampleRate=fs, window="hann")
    assert (rec["index"][idx] == ref["peaks"]["index"]).all()
    assert np.abs(rec["amplitude"][idx] - ref["peaks"]["amplitude"]).max() <= 1e-12


def test_c3_full_size_stft_against_oracle_subset(ctx):
    """BASELINE config C3 at its full size: 10 min at 48 kHz (28,800,000 samples), N=4096, hop 1024, Hann, amplitude +
    phase in fp64 = 28,122 overlapping frames read straight from the stream (no frame matrix).  Frame count and the
    frames' positions are checked through a strided subset against the oracle; Parseval holds for every frame."""
    import torch
    from pragma_dsp_b200 import stft
    n, hop, total = 4096, 1024, 28_800_000
    rng = np.random.default_rng(3)
    t = np.arange(total)
    sig = np.zeros(total)
    for k in range(4):  # slowly drifting tones: every frame has a different spectrum
        f0, f1 = rng.uniform(200, 8000, 2)
        sig += rng.uniform(0.2, 1.0) * np.sin(2 * np.pi * (f0 + (f1 - f0) * t / (2 * total)) * t / 48000.0 + rng.uniform(0, 6.28))
    got = stft(sig, fftSize=n, hopSize=hop, window="hann", sampleRate=48000.0, outputs=("amplitude", "phase", "peak"))
    frames = (total - n) // hop + 1
    assert frames == 28122 and got["amplitude"].shape == (frames, n // 2 + 1) and got["phase"].shape == (frames, n // 2 + 1)
    idx = np.arange(0, frames, 997)
    sub = np.stack([sig[i * hop:i * hop + n] for i in idx])
    ref = oracle.spectrum_batch(sub, fftSize=n, sampleRate=48000.0, window="hann")
    assert np.abs(got["amplitude"][idx] - ref["amplitude"]).max() <= 1e-12
    strong = ref["amplitude"] > 1e-6  # phase is compared where |X| is not noise, wrap-aware (test/reallife/signals.test.ts:34-49)
    dphi = np.angle(np.exp(1j * (got["phase"][idx] - ref["phase"])))
    assert np.abs(dphi[strong]).max() <= 1e-8
    assert (got["peaks"]["index"][idx] == ref["peaks"]["index"]).all()
    # Parseval on the windowed frames: sum over the one-sided bins of (scaled amplitude)^2 recovers 2/N * sum (w*x)^2
    w = oracle.createWindow("hann", n)
    a = got["amplitude"][idx]
    lhs = (a[:, 1:-1] ** 2).sum(1) / 2 + a[:, 0] ** 2 + a[:, -1] ** 2
    rhs = ((sub * w) ** 2).sum(1) / n
    assert np.abs(lhs - rhs).max() <= 1e-10 * rhs.max()


@pytest.mark.parametrize("log2n", [16, 18, 20, 24])
def test_large_fft_pass_generations_agree(ctx, log2n):
    """K2 has three generations of the pass kernel (per-thread / TMA loads; TMA loads + stores, two CTAs per SM; pipeline
    passes with the split-plane exchange) and a fused form of the middle + last pass (one persistent launch, tiles handed
    over through per-group counters).  Forced one at a time through the tunables, all give the same transform."""
    from pragma_dsp_b200.core import ComplexArray, Radix2Fft
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    re, im = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    ref = np.fft.fft(re + 1j * im)
    outs = []
    try:
        for pipe, v2, fused in (("0", "0", None), ("0", "1", None), ("1", "0", None), (None, None, None), (None, None, "1"), ("0", "1", "1")):
            if fused and log2n < 22:
                continue  # the fused middle + last pass exists for three-pass transforms
            ctx.tune("big_pipe", pipe)
            ctx.tune("big_v2", v2)
            ctx.tune("big_fused", fused)
            out = Radix2Fft(n).forwardComplex(ComplexArray(re, im))
            z = out.real + 1j * out.imag
            assert np.linalg.norm(z - ref) / np.linalg.norm(ref) <= 1e-12 * log2n
            outs.append(z)
            back = Radix2Fft(n).inverse(out)
            assert np.abs(back.real - re).max() <= 1e-12 and np.abs(back.imag - im).max() <= 1e-12
    finally:
        ctx.tune("big_pipe", None)
        ctx.tune("big_v2", None)
        ctx.tune("big_fused", None)
    for z in outs[1:]:  # same butterflies, same twiddles: the generations differ in data movement only
        assert np.linalg.norm(z - outs[0]) / np.linalg.norm(outs[0]) <= 1e-15


@pytest.mark.parametrize("n", [64, 1024, 4096])
def test_edge_frames_in_one_batch_match_the_oracle(ctx, n):
    """Frames that take the kernels' rare paths, mixed into one batch so that a warp / CTA sees them side by side with ordinary
    frames: magnitudes whose squares underflow or overflow (the exponent tracker reruns the epilogue with hypot), NaN and
    Inf samples (findPeak: NaN never wins, src/public/spectrum.ts:74-105), silence, pure DC, a Nyquist tone, exact ties."""
    from pragma_dsp_b200 import spectrum_batch
    rng = np.random.default_rng(n)
    t = np.arange(n)
    x = multitone(rng, 40, n)
    x[3] *= 1e-170
    x[4] *= 1e170
    x[5, 7] = np.nan
    x[6, 11] = np.inf
    x[7] = 0.0
    x[8] = 0.75
    x[9] = np.where(t % 2 == 0, 1.0, -1.0)
    x[10] = np.sin(2 * np.pi * 5 * t / n) + np.sin(2 * np.pi * 9 * t / n)  # two equal peaks: the lower bin wins
    x[11] *= 1e-300  # subnormal products
    for window in ("rect", "hann"):
        got = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window=window, precision="f64", context=ctx)
        ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window=window)
        gi, ri = got["peaks"]["index"], ref["peaks"]["index"]
        finite = np.isfinite(ref["amplitude"]).all(axis=1)
        sure = finite.copy()
        sure[10] = False  # bins 5 and 9 tie to rounding noise: which of the two wins depends on the last bit of each FFT
        assert (gi[sure] == ri[sure]).all(), (window, np.nonzero(gi != ri))
        assert gi[10] in (5, 9) and ri[10] in (5, 9)
        assert gi[5] == ri[5] == 0  # an all-NaN spectrum: no comparison is ever true, index 0 (NaN never wins)
        # an Inf sample leaves a mix of NaN and Inf bins whose positions depend on the factorisation (hypot(NaN, Inf) = Inf):
        # unpinned by the reference; only "nothing finite comes out" is common ground
        assert np.isnan(got["amplitude"][5]).all() and not np.isfinite(got["amplitude"][6]).any()
        scale = np.maximum(np.abs(ref["amplitude"]).max(axis=1, keepdims=True), 1e-300)
        err = np.abs(got["amplitude"][finite] - ref["amplitude"][finite]) / scale[finite]
        assert err.max() <= 1e-12, (window, err.max(), np.unravel_index(err.argmax(), err.shape))
        assert (got["amplitude"][7] == 0).all() and got["peaks"]["index"][7] == 0


def test_legacy_default_stream_handle(ctx):
    """A caller that works on CUDA's legacy default stream passes cudaStreamLegacy ((void*)1): the launch is then ordered
    behind the producer of its inputs on that stream without any synchronisation (handle 0 would select the context's own,
    non-blocking stream - include/pragma_b200.h)."""
    import torch
    from pragma_dsp_b200._lib import F64, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib
    L = lib()
    n, batch = 1024, 50_000
    rng = np.random.default_rng(21)
    host = multitone(rng, 64, n)
    x = torch.from_numpy(host).cuda().repeat(batch // 64 + 1, 1)[:batch].contiguous()
    big = torch.empty((batch, n), dtype=torch.float64, device="cuda")
    amp = torch.empty((batch, n // 2 + 1), dtype=torch.float64, device="cuda")
    pk = torch.zeros((batch, 32), dtype=torch.uint8, device="cuda")
    d = SpectrumDesc(sample_dtype=F64, frame_len=n, hop=n, batch=batch, window=WINDOWS["hann"], sides=SIDES["one"],
                     sample_rate=48000.0, raw_magnitude=0, fft_shift=0)
    plan = ctx.plan(n, F64)
    torch.cuda.synchronize()
    for _ in range(4):  # a queue of default-stream work the library's launch has to wait for
        big.copy_(x).mul_(1.0)
    amp.zero_()
    check(L.pdsp_spectrum_dev(plan, C.byref(d), C.c_void_p(big.data_ptr()), C.c_void_p(amp.data_ptr()), None,
                              C.c_void_p(pk.data_ptr()), C.c_void_p(1)))
    torch.cuda.synchronize()
    ref = oracle.spectrum_batch(host, fftSize=n, sampleRate=48000.0, window="hann")
    got = amp.cpu().numpy()
    assert np.abs(got[:64] - ref["amplitude"]).max() <= 1e-12
    assert np.array_equal(got[:64], got[batch - batch % 64 - 64:batch - batch % 64])  # the same 64 frames further down the batch
    assert (pk.cpu().numpy().view(PEAK_F64).reshape(-1)["index"][:64] == ref["peaks"]["index"]).all()
