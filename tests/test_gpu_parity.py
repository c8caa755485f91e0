"""Parity of the CUDA path against the CPU oracle on seeded inputs, through the C-ABI
(host entry points and device-resident entry points).  Run on the B200 box: pytest -m gpu.

Tolerances (BASELINE.json north_star): bin/peak indices exact; complex outputs within relative
L2 1e-12*log2(N) for fp64 and 1e-5 for fp32."""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def tol(n, prec):
    return 1e-12 * max(1.0, np.log2(n)) if prec == "f64" else 1e-5


def rel_l2(a, b):
    s = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (s if s > 0 else 1.0)


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


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_forward_real_all_sizes(n):
    from pragma_dsp_b200.core import Radix2Fft
    rng = np.random.default_rng(n)
    batch = 37 if n <= 4096 else 5
    x = rng.standard_normal((batch, n))
    re, im = Radix2Fft(n).forward_batch(x)
    rre, rim = oracle.FFT(n).forward(x)
    for f in range(batch):
        assert rel_l2(re[f] + 1j * im[f], rre[f] + 1j * rim[f]) <= tol(n, "f64"), (n, f)
    # Float32Array input is widened exactly (README.md:11)
    re32, im32 = Radix2Fft(n).forward_batch(x.astype(np.float32))
    rre, rim = oracle.FFT(n).forward(x.astype(np.float32).astype(np.float64))
    assert rel_l2(re32 + 1j * im32, rre + 1j * rim) <= tol(n, "f64")


@pytest.mark.parametrize("n", [1, 2, 4, 8, 32, 64, 512, 1024, 4096, 8192])
@pytest.mark.parametrize("inverse", [False, True])
def test_complex_all_sizes(n, inverse):
    from pragma_dsp_b200.core import Radix2Fft
    rng = np.random.default_rng(1000 + n)
    batch = 19 if n <= 1024 else 4
    re, im = rng.standard_normal((batch, n)), rng.standard_normal((batch, n))
    ore, oim = Radix2Fft(n).complex_batch(re, im, inverse=inverse)
    plan = oracle.FFT(n)
    rre, rim = (plan.inverse if inverse else plan.forwardComplex)(re, im)
    for f in range(batch):
        assert rel_l2(ore[f] + 1j * oim[f], rre[f] + 1j * rim[f]) <= tol(n, "f64"), (n, f)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("n,window,sides", [(1024, "hann", "one"), (1024, "rect", "two"), (4096, "hann", "one"),
                                            (256, "blackman", "one"), (64, "hamming", "two"), (2048, "hamming", "one"),
                                            (16, "hann", "one"), (2, "rect", "one"), (1, "rect", "one"), (8192, "hann", "one")])
def test_spectrum_batch_vs_oracle(prec, n, window, sides):
    from pragma_dsp_b200 import spectrum_batch
    rng = np.random.default_rng(n * 7 + len(window))
    batch = 300 if n <= 1024 else 40
    x = multitone(rng, batch, n) if n >= 64 else rng.standard_normal((batch, n))
    xs = x.astype(np.float32) if prec == "f32" else x
    got = spectrum_batch(xs, sampleRate=48000.0, fftSize=n, window=window, sides=sides, precision=prec)
    ref = oracle.spectrum_batch(xs, fftSize=n, sampleRate=48000.0, window=window, sides=sides)
    if sides == "one":
        assert (got["peaks"]["index"] == ref["peaks"]["index"]).all()
    else:  # the reference picks k or N-k on 1-ulp noise
        gi, ri = got["peaks"]["index"], ref["peaks"]["index"]
        assert ((gi == ri) | (gi == (n - ri) % n)).all()
    atol = 1e-12 if prec == "f64" else 2e-6
    scale = np.maximum(1.0, np.abs(ref["amplitude"]).max(axis=1, keepdims=True))
    assert (np.abs(got["amplitude"] - ref["amplitude"]) <= 10 * atol * scale).all()
    for f in range(0, batch, 7):
        assert rel_l2(got["amplitude"][f].astype(np.float64), ref["amplitude"][f]) <= tol(n, prec)
    mask = ref["amplitude"] > (1e-6 if prec == "f64" else 1e-3) * scale
    d = np.abs(got["phase"] - ref["phase"])
    d = np.minimum(d, np.abs(d - 2 * np.pi))
    assert d[mask].max(initial=0) <= (1e-9 if prec == "f64" else 2e-3)
    pa = got["peaks"]["amplitude"]
    assert (np.abs(pa - ref["peaks"]["amplitude"]) <= 10 * atol * scale[:, 0]).all()
    assert np.allclose(got["peaks"]["frequency"], got["peaks"]["index"] * (48000.0 / n), rtol=1e-6 if prec == "f32" else 1e-15)


def test_frame_addressing_host_api():
    """buildFrame zero-pad / truncate (src/public/spectrum.ts:36-43), STFT hop views, odd alignments."""
    from pragma_dsp_b200 import spectrum_batch
    rng = np.random.default_rng(3)
    sig = rng.standard_normal(40000).astype(np.float32)
    for n, frame_len, hop, batch in [(256, 100, 37, 900), (4096, 4096, 1024, 30), (128, 300, 301, 100), (64, 0, 5, 3),
                                     (512, 511, 1, 700), (1024, 1024, 1, 33)]:
        got = spectrum_batch(sig, fftSize=n, frameLen=frame_len, hop=hop, batch=batch, sampleRate=8000.0, window="hann")
        ref = oracle.spectrum_batch(sig, fftSize=n, frameLen=frame_len, hop=hop, batch=batch, sampleRate=8000.0, window="hann")
        assert np.abs(got["amplitude"] - ref["amplitude"]).max() <= 1e-13
        assert (got["peaks"]["index"] == ref["peaks"]["index"]).all()
    # f64 samples at an address that is 8- but not 16-byte aligned (scalar-load path)
    buf = rng.standard_normal(1024 * 8 + 1)
    got = spectrum_batch(buf[1:], fftSize=1024, frameLen=1024, hop=1024, batch=8, window="blackman")
    ref = oracle.spectrum_batch(buf[1:], fftSize=1024, frameLen=1024, hop=1024, batch=8, window="blackman")
    assert np.abs(got["amplitude"] - ref["amplitude"]).max() <= 1e-13


def test_chunked_pipeline_large_batch_pageable_and_pinned(ctx):
    """Host entry point with a batch that spans many staging chunks; pageable and pinned buffers agree."""
    import torch
    from pragma_dsp_b200 import spectrum_batch
    rng = np.random.default_rng(5)
    n, batch = 1024, 20000  # ~82 MB in, 3+ chunks
    x = multitone(rng, batch, n, np.float32)
    got = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann", precision="f32", outputs=("amplitude", "peak"))
    ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window="hann", want_phase=False, threads=8)
    assert (got["peaks"]["index"] == ref["peaks"]["index"]).all()
    assert np.abs(got["amplitude"] - ref["amplitude"]).max() <= 2e-5
    px = torch.from_numpy(x).pin_memory().numpy()
    got2 = spectrum_batch(px, sampleRate=48000.0, fftSize=n, window="hann", precision="f32", outputs=("amplitude", "peak"))
    assert (got2["amplitude"] == got["amplitude"]).all() and (got2["peaks"] == got["peaks"]).all()


def test_device_resident_entry_points(ctx):
    """pdsp_spectrum_dev / pdsp_fft_*_dev on torch-owned device memory and torch's stream."""
    import torch
    from pragma_dsp_b200._lib import F32, F64, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib
    L = lib()
    rng = np.random.default_rng(9)
    n, batch = 1024, 513
    x = multitone(rng, batch, n)
    dx = torch.from_numpy(x).cuda()
    amp = torch.empty((batch, n // 2 + 1), dtype=torch.float64, device="cuda")
    pk = torch.zeros((batch, 32), dtype=torch.uint8, device="cuda")
    plan = ctx.plan(n, F64)
    d = SpectrumDesc(sample_dtype=F64, frame_len=n, hop=n, batch=batch, window=WINDOWS["hann"], sides=SIDES["one"],
                     sample_rate=48000.0, raw_magnitude=0)
    # torch's current stream is the legacy default stream (handle 0 = NULL = "the context's stream" to the library, a
    # non-blocking stream that is not ordered behind torch's work): finish the uploads and fills first
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream
    check(L.pdsp_spectrum_dev(plan, C.byref(d), C.c_void_p(dx.data_ptr()), C.c_void_p(amp.data_ptr()), None,
                              C.c_void_p(pk.data_ptr()), C.c_void_p(st)))
    torch.cuda.synchronize()
    ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window="hann")
    assert np.abs(amp.cpu().numpy() - ref["amplitude"]).max() <= 1e-12
    peaks = pk.cpu().numpy().view(PEAK_F64).reshape(-1)
    assert (peaks["index"] == ref["peaks"]["index"]).all()
    # real forward, one-sided planes, then complex inverse of the full spectrum = the signal
    ore = torch.empty((batch, n), dtype=torch.float64, device="cuda")
    oim = torch.empty((batch, n), dtype=torch.float64, device="cuda")
    check(L.pdsp_fft_forward_real_dev(plan, C.c_void_p(dx.data_ptr()), F64, batch, C.c_void_p(ore.data_ptr()),
                                      C.c_void_p(oim.data_ptr()), 1, C.c_void_p(st)))
    bre, bim = torch.empty_like(ore), torch.empty_like(ore)
    check(L.pdsp_fft_complex_dev(plan, C.c_void_p(ore.data_ptr()), C.c_void_p(oim.data_ptr()), batch,
                                 C.c_void_p(bre.data_ptr()), C.c_void_p(bim.data_ptr()), 1, C.c_void_p(st)))
    torch.cuda.synchronize()
    assert (bre - dx).abs().max().item() <= 1e-13 and bim.abs().max().item() <= 1e-13


@pytest.mark.parametrize("prec,frames", [("f32", 65536), ("f64", 65536)])
def test_full_size_properties(ctx, prec, frames):
    """BASELINE-size batch (65,536 x 1024), checked through size-independent properties:
    Parseval, linearity, inverse round trip, and oracle agreement on a strided subset."""
    import torch
    from pragma_dsp_b200._lib import F32, F64, PEAK_F32, PEAK_F64, SIDES, WINDOWS, SpectrumDesc, check, lib
    L = lib()
    n = 1024
    P = F64 if prec == "f64" else F32
    tdt = torch.float64 if prec == "f64" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(1337)
    x = torch.randn((frames, n), generator=g, device="cuda", dtype=torch.float64).to(tdt)
    y = torch.randn((frames, n), generator=g, device="cuda", dtype=torch.float64).to(tdt)
    plan = ctx.plan(n, P)
    torch.cuda.synchronize()  # handle 0 below = the context's own non-blocking stream: not ordered behind torch's generators
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fwd(t):
        re = torch.empty((frames, n), dtype=tdt, device="cuda")
        im = torch.empty((frames, n), dtype=tdt, device="cuda")
        check(L.pdsp_fft_forward_real_dev(plan, C.c_void_p(t.data_ptr()), P, frames, C.c_void_p(re.data_ptr()),
                                          C.c_void_p(im.data_ptr()), 1, st))
        return re, im

    xr, xi = fwd(x)
    yr, yi = fwd(y)
    sr, si = fwd(x + 2 * y)
    torch.cuda.synchronize()
    eps = 1e-12 if prec == "f64" else 2e-5
    # Parseval: sum |X|^2 = N * sum x^2, per frame
    e_t = (x.double() ** 2).sum(1) * n
    e_f = (xr.double() ** 2 + xi.double() ** 2).sum(1)
    assert ((e_f - e_t).abs() / e_t).max().item() <= 50 * eps
    # linearity
    num = ((sr - (xr + 2 * yr)).double() ** 2 + (si - (xi + 2 * yi)).double() ** 2).sum(1).sqrt()
    den = (sr.double() ** 2 + si.double() ** 2).sum(1).sqrt()
    assert (num / den).max().item() <= 50 * eps
    # Hermitian symmetry is exact by construction
    assert (xr[:, 1:n // 2] == xr[:, n // 2 + 1:].flip(1)).all() and (xi[:, 1:n // 2] == -xi[:, n // 2 + 1:].flip(1)).all()
    # inverse round trip
    br, bi = torch.empty_like(xr), torch.empty_like(xr)
    check(L.pdsp_fft_complex_dev(plan, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), frames,
                                 C.c_void_p(br.data_ptr()), C.c_void_p(bi.data_ptr()), 1, st))
    torch.cuda.synchronize()
    assert (br - x).abs().max().item() <= 100 * eps and bi.abs().max().item() <= 100 * eps
    # oracle agreement on a strided subset, including the peaks
    idx = torch.arange(0, frames, 257, device="cuda")
    sub = x[idx].cpu().numpy()
    rre, rim = oracle.FFT(n).forward(sub.astype(np.float64))
    got = (xr[idx].cpu().numpy().astype(np.float64) + 1j * xi[idx].cpu().numpy().astype(np.float64))
    for f in range(len(idx)):
        assert rel_l2(got[f], rre[f] + 1j * rim[f]) <= tol(n, prec)


@pytest.mark.parametrize("log2n", [14, 15, 17, 20, 21, 24])
def test_large_multipass_fft(ctx, log2n):
    """BASELINE config C4 (N = 2^20, 2^24 complex fp64) and neighbours: the multi-pass path against
    numpy's pocketfft (the reference's own golden generator, scripts/gen_fixtures.py:11-13) and, at 2^20,
    against the oracle.  Tolerance: relative L2 1e-12*log2(N) (north_star)."""
    from pragma_dsp_b200.core import ComplexArray, Radix2Fft
    n = 1 << log2n
    rng = np.random.default_rng(1337)
    re, im = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    fft = Radix2Fft(n)
    out = fft.forwardComplex(ComplexArray(re, im))
    ref = np.fft.fft(re + 1j * im)
    tol_ = 1e-12 * log2n
    assert rel_l2(out.real + 1j * out.imag, ref) <= tol_
    if log2n == 20:
        rre, rim = oracle.FFT(n).forwardComplex(re, im)
        assert rel_l2(out.real + 1j * out.imag, rre[0] + 1j * rim[0] if rre.ndim > 1 else rre + 1j * rim) <= tol_
    back = fft.inverse(out)
    assert np.abs(back.real - re).max() <= 1e-12 and np.abs(back.imag - im).max() <= 1e-12
    if log2n in (15, 20):
        x = rng.standard_normal(n)
        r = fft.forward(x)  # real input through the same path
        assert rel_l2(r.real + 1j * r.imag, np.fft.fft(x)) <= tol_
    # exact-zero rule survives the multi-pass path
    z = fft.forwardComplex(ComplexArray(np.zeros(n), np.zeros(n)))
    assert (z.real == 0).all() and (z.imag == 0).all()


def test_large_multipass_fft_fp32_device(ctx):
    import torch
    from pragma_dsp_b200._lib import F32, check, lib
    L = lib()
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(7)
    re = torch.rand(n, generator=g, device="cuda") * 2 - 1
    im = torch.rand(n, generator=g, device="cuda") * 2 - 1
    ore, oim = torch.empty_like(re), torch.empty_like(re)
    plan = ctx.plan(n, F32)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())  # ordered behind the generators / fills on torch's stream
    check(L.pdsp_fft_complex_dev(plan, C.c_void_p(re.data_ptr()), C.c_void_p(im.data_ptr()), 1, C.c_void_p(ore.data_ptr()),
                                 C.c_void_p(oim.data_ptr()), 0, C.c_void_p(st.cuda_stream)))
    st.synchronize()
    ref = np.fft.fft(re.cpu().numpy().astype(np.float64) + 1j * im.cpu().numpy().astype(np.float64))
    got = ore.cpu().numpy().astype(np.float64) + 1j * oim.cpu().numpy().astype(np.float64)
    assert rel_l2(got, ref) <= 1e-5


def test_fused_peer_scatter_single_gpu(ctx):
    """pdsp_spectrum_dev_gather with the rank's own buffer (and a second local buffer standing in for a
    peer): every finished peak record lands at record_offset + f in each target."""
    import torch
    from pragma_dsp_b200._lib import F32, PEAK_F32, SIDES, WINDOWS, SpectrumDesc, check, lib
    L = lib()
    rng = np.random.default_rng(11)
    n, batch, offset = 1024, 777, 100
    x = multitone(rng, batch, n, np.float32)
    dx = torch.from_numpy(x).cuda()
    local = torch.zeros((batch, 16), dtype=torch.uint8, device="cuda")
    tgt = [torch.zeros((offset + batch + 5, 16), dtype=torch.uint8, device="cuda") for _ in range(2)]
    peers = (C.c_void_p * 8)()
    peers[0], peers[1] = tgt[0].data_ptr(), tgt[1].data_ptr()
    plan = ctx.plan(n, F32)
    d = SpectrumDesc(sample_dtype=F32, frame_len=n, hop=n, batch=batch, window=WINDOWS["hann"], sides=SIDES["one"],
                     sample_rate=48000.0, raw_magnitude=0)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())  # ordered behind the generators / fills on torch's stream
    check(L.pdsp_spectrum_dev_gather(plan, C.byref(d), C.c_void_p(dx.data_ptr()), None, None, C.c_void_p(local.data_ptr()),
                                     peers, 2, offset, C.c_void_p(st.cuda_stream)))
    st.synchronize()
    ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window="hann")
    loc = local.cpu().numpy().view(PEAK_F32).reshape(-1)
    assert (loc["index"] == ref["peaks"]["index"]).all()
    for t_ in tgt:
        rec = t_.cpu().numpy().view(PEAK_F32).reshape(-1)
        assert (rec[offset:offset + batch] == loc).all()
        assert (rec[:offset]["index"] == 0).all() and (rec[offset + batch:]["index"] == 0).all()


@pytest.mark.parametrize("tma,interleave,prefetch", [("1", "1", "1"), ("0", "1", "1"), ("1", "0", "1"), ("0", "0", "0"),
                                                     ("1", "1", "0")])
def test_large_fft_tma_and_plain_paths_agree(ctx, tma, interleave, prefetch):
    """K2's strided passes have two implementations: TMA tile loads (cp.async.bulk.tensor + mbarrier) and
    per-thread coalesced loads (tunable big_tma=0), an interleaved or planar work buffer between the passes
    (big_interleave) and an optional L2 prefetch of the next tile.  All must give the same transform."""
    from pragma_dsp_b200.core import ComplexArray, Radix2Fft
    ctx.tune("big_tma", tma)
    ctx.tune("big_interleave", interleave)
    ctx.tune("big_prefetch", prefetch)
    try:
        for log2n in (14, 18, 22):
            n = 1 << log2n
            rng = np.random.default_rng(log2n)
            re, im = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
            out = Radix2Fft(n).forwardComplex(ComplexArray(re, im))
            assert rel_l2(out.real + 1j * out.imag, np.fft.fft(re + 1j * im)) <= 1e-12 * log2n
    finally:
        for k in ("big_tma", "big_interleave", "big_prefetch"):
            ctx.tune(k, None)


def test_fast_atan2_accuracy_and_special_values(ctx):
    """The phase rows use a hand-written atan2 (one division + odd minimax polynomial).  Check the device
    implementation through phase() against libm over all quadrants, ratios near the octant folds, huge /
    tiny magnitudes (libdevice escape path) and the signed-zero cases Math.atan2 defines."""
    from pragma_dsp_b200.core import ComplexArray
    from pragma_dsp_b200.xform import phase
    rng = np.random.default_rng(2)
    n = 1 << 18
    re = rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 6, n)
    im = rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 6, n)
    # ratios around tan(pi/8), 1 and the axes
    k = 4096
    ang = np.concatenate([np.linspace(-np.pi, np.pi, k), np.pi / 8 + np.linspace(-1e-9, 1e-9, k), np.pi / 4 + np.linspace(-1e-9, 1e-9, k)])
    re[:ang.size], im[:ang.size] = np.cos(ang) * 3.7, np.sin(ang) * 3.7
    re[20000:20004], im[20000:20004] = [1e300, -1e300, 1e-300, -1e-300], [1e300, 1e-300, 1e-300, 1e300]
    got = phase(ComplexArray(re, im))
    ref = np.arctan2(im, re)
    err = np.abs(got - ref)
    assert err.max() <= 4 * np.finfo(np.float64).eps * np.pi, err.max()
    sp_re = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 0.0, 0.0, -2.0, 2.0])
    sp_im = np.array([0.0, 0.0, -0.0, -0.0, 0.0, 0.0, 1.0, -1.0, -0.0, -0.0])
    got = phase(ComplexArray(sp_re, sp_im))
    ref = np.arctan2(sp_im, sp_re)
    assert (got == ref).all() and (np.signbit(got) == np.signbit(ref)).all(), (got, ref)


def test_device_resident_fft_convolution(ctx):
    """forward -> frequency-domain product -> inverse without leaving the device
    (the fluent layer's convolution, test/fluent/chain.test.ts:287-316, at N = 4096 x 64 frames)."""
    import torch
    from pragma_dsp_b200._lib import F64, check, lib
    L = lib()
    n, batch = 4096, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((batch, n), generator=g, device="cuda", dtype=torch.float64)
    b = torch.randn((batch, n), generator=g, device="cuda", dtype=torch.float64)
    plan = ctx.plan(n, F64)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())  # ordered behind the generators / fills on torch's stream
    s = C.c_void_p(st.cuda_stream)
    vp = lambda t_: C.c_void_p(t_.data_ptr())  # noqa: E731
    bufs = [torch.empty((batch, n), dtype=torch.float64, device="cuda") for _ in range(8)]
    are, aim, bre, bim, pre, pim, ore, oim = bufs
    check(L.pdsp_fft_forward_real_dev(plan, vp(a), F64, batch, vp(are), vp(aim), 1, s))
    check(L.pdsp_fft_forward_real_dev(plan, vp(b), F64, batch, vp(bre), vp(bim), 1, s))
    check(L.pdsp_complex_mul_dev(ctx.h, F64, vp(are), vp(aim), vp(bre), vp(bim), 0, 1.0, batch * n, vp(pre), vp(pim), s))
    check(L.pdsp_fft_complex_dev(plan, vp(pre), vp(pim), batch, vp(ore), vp(oim), 1, s))
    st.synchronize()
    ref = np.fft.ifft(np.fft.fft(a.cpu().numpy(), axis=1) * np.fft.fft(b.cpu().numpy(), axis=1), axis=1)
    assert np.abs(ore.cpu().numpy() - ref.real).max() <= 1e-10 and np.abs(oim.cpu().numpy()).max() <= 1e-10
    # correlation: conj on the second operand, scale folded in
    check(L.pdsp_complex_mul_dev(ctx.h, F64, vp(are), vp(aim), vp(bre), vp(bim), 1, 0.5, batch * n, vp(pre), vp(pim), s))
    st.synchronize()
    refp = 0.5 * np.fft.fft(a.cpu().numpy(), axis=1) * np.conj(np.fft.fft(b.cpu().numpy(), axis=1))
    assert np.abs(pre.cpu().numpy() - refp.real).max() <= 1e-9 and np.abs(pim.cpu().numpy() - refp.imag).max() <= 1e-9


@pytest.mark.parametrize("log2n,precision", [(15, "f64"), (17, "f64"), (20, "f64"), (16, "f32")])
def test_large_spectrum_matches_oracle(ctx, log2n, precision):
    """spectrum() for FFT sizes beyond the fused single-CTA kernel (N > 16384): buildFrame + window kernel,
    multi-pass transform, epilogue kernel with findPeak.  Amplitude, phase and the peak record against the oracle,
    including a zero-padded frame length and a silent frame."""
    from pragma_dsp_b200 import spectrum_batch
    n = 1 << log2n
    batch = 3 if log2n <= 17 else 2
    rng = np.random.default_rng(log2n)
    flen = n - 1234
    dt = np.float64 if precision == "f64" else np.float32
    x = multitone(rng, batch, flen).astype(dt)
    x[batch - 1] = 0.0
    got = spectrum_batch(x, sampleRate=96000.0, fftSize=n, window="blackman", precision=precision)
    ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=96000.0, window="blackman", threads=8)
    atol = 1e-12 if precision == "f64" else 3e-6
    assert got["amplitude"].shape == (batch, n // 2 + 1)
    assert np.abs(got["amplitude"] - ref["amplitude"]).max() <= atol
    assert (got["peaks"]["index"] == ref["peaks"]["index"]).all()
    assert np.abs(got["peaks"]["amplitude"] - ref["peaks"]["amplitude"]).max() <= atol
    assert np.abs(got["peaks"]["frequency"] - ref["peaks"]["frequency"]).max() <= (1e-9 if precision == "f64" else 1e-2)
    strong = ref["amplitude"] > (1e-6 if precision == "f64" else 1e-3)
    d = np.abs(got["phase"] - ref["phase"])
    assert np.minimum(d, np.abs(d - 2 * np.pi))[strong].max() <= (1e-7 if precision == "f64" else 2e-2)


@pytest.mark.parametrize("n", [2, 4, 64, 1024, 8192, 1 << 15])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_fused_fftshift_equals_shift_after(ctx, n, precision):
    """desc.fft_shift stores two-sided amplitude / phase rows already fftShift-ed (fourier.ts:122-134);
    bit-identical to shifting the unshifted rows afterwards; the peak record is unaffected."""
    from pragma_dsp_b200 import spectrum_batch
    rng = np.random.default_rng(n)
    x = (multitone(rng, 5, n) if n >= 64 else rng.standard_normal((5, n))).astype(np.float64 if precision == "f64" else np.float32)
    plain = spectrum_batch(x, sampleRate=8000.0, fftSize=n, window="hamming", sides="two", precision=precision)
    fused = spectrum_batch(x, sampleRate=8000.0, fftSize=n, window="hamming", sides="two", precision=precision, shift=True)
    assert np.array_equal(fused["amplitude"], np.fft.fftshift(plain["amplitude"], axes=1))
    assert np.array_equal(fused["phase"], np.fft.fftshift(plain["phase"], axes=1))
    assert (fused["peaks"] == plain["peaks"]).all()


def test_ingestion_ring_matches_batched_call(ctx):
    """10,000 frames pushed through the ingestion ring in uneven blocks come back in order and
    bit-identical to one pdsp_spectrum call; fp32 frames, fp32 plan; back-pressure and the final partial chunk."""
    from pragma_dsp_b200 import IngestRing, spectrum_batch
    rng = np.random.default_rng(5)
    n, total = 1024, 10000
    x = multitone(rng, total, n, np.float32)
    ref = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann", precision="f32", outputs=("amplitude", "peak"))
    amp, pk = [], []
    with IngestRing(n, sampleRate=48000.0, window="hann", precision="f32", sample_dtype=np.float32,
                    outputs=("amplitude", "peak"), framesPerChunk=1024, depth=3) as ring:
        pushed = 0
        while pushed < total:
            k = ring.push(x[pushed:pushed + 777])
            pushed += k
            if k == 0:
                out = ring.pop(1500)
                assert out["count"] > 0 and out["phase"] is None
                amp.append(out["amplitude"]), pk.append(out["peaks"])
        ring.flush()
        while True:
            out = ring.pop(4096)
            if out["count"] == 0:
                break
            amp.append(out["amplitude"]), pk.append(out["peaks"])
    amp, pk = np.concatenate(amp), np.concatenate(pk)
    assert amp.shape == ref["amplitude"].shape
    assert np.array_equal(amp, ref["amplitude"]) and (pk == ref["peaks"]).all()


def test_randomised_shapes_against_oracle_gpu(ctx):
    """Seeded random walk over the spectrum() parameter space on the device (every size class, both plan precisions,
    zero-padding / truncation, overlapping and odd hops, both sides, all windows, output subsets, fused shift)."""
    from pragma_dsp_b200 import spectrum_batch
    import os
    rng = np.random.default_rng(int(os.environ.get("PDSP_RANDOM_SEED", "77")))
    sizes = [1, 2, 4, 8, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768]
    for case in range(int(os.environ.get("PDSP_RANDOM_CASES", "120"))):
        n = int(rng.choice(sizes))
        batch = int(rng.integers(1, 40 if n <= 4096 else 4))
        frame_len = int(rng.choice([n, max(1, n - 1), max(1, n // 2 + 1), n + 3, 1]))
        hop = int(rng.choice([frame_len, max(1, frame_len // 2), frame_len + 5, max(1, frame_len - 1)]))
        window = str(rng.choice(["rect", "hann", "hamming", "blackman"]))
        sides = str(rng.choice(["one", "two"]))
        sdt = np.float64 if rng.integers(0, 2) else np.float32
        precision = "f64" if rng.integers(0, 3) else "f32"
        outputs = [("amplitude", "phase", "peak"), ("amplitude",), ("peak",), ("amplitude", "peak"), ("phase",)][int(rng.integers(0, 5))]
        shift = bool(sides == "two" and rng.integers(0, 2))
        total = (batch - 1) * hop + frame_len
        x = rng.standard_normal(total).astype(sdt)
        if n >= 16:
            k0 = int(rng.integers(1, n // 2))
            x += (3.0 * np.sin(2 * np.pi * k0 * np.arange(total) / n)).astype(sdt)
        tag = (case, n, batch, frame_len, hop, window, sides, sdt.__name__, precision, outputs, shift)
        got = spectrum_batch(x, sampleRate=8000.0, fftSize=n, window=window, sides=sides, frameLen=frame_len, hop=hop,
                             batch=batch, outputs=outputs, shift=shift, precision=precision)
        ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=8000.0, window=window, sides=sides, frameLen=frame_len,
                                    hop=hop, batch=batch, threads=8)
        ramp, rph = ref["amplitude"], ref["phase"]
        if shift:
            ramp, rph = np.fft.fftshift(ramp, axes=1), np.fft.fftshift(rph, axes=1)
        scale = max(1.0, float(np.abs(ramp).max()))
        atol = (1e-12 if precision == "f64" else 2e-6 * np.log2(max(2, n))) * scale
        if "amplitude" in outputs:
            assert np.abs(got["amplitude"] - ramp).max() <= atol, tag
        if "phase" in outputs:
            strong = ramp > (1e-6 if precision == "f64" else 1e-2) * scale
            d = np.abs(got["phase"].astype(np.float64) - rph)
            if strong.any():
                assert np.minimum(d, np.abs(d - 2 * np.pi))[strong].max() <= (1e-8 if precision == "f64" else 5e-3), tag
        if "peak" in outputs:
            gi, ri = got["peaks"]["index"], ref["peaks"]["index"]
            for f in np.nonzero(gi != ri)[0]:  # only last-bit ties (noise bins, mirror bins) may differ
                a = ref["amplitude"][f]
                if precision == "f32" and min(a[gi[f]], a[ri[f]]) * n < 1e-18:
                    continue  # a bin whose |X|^2 is below the fp32 normal range flushes to 0 (sqrt.approx.ftz) and cannot win
                assert abs(a[gi[f]] - a[ri[f]]) <= (1e-12 if precision == "f64" else 1e-5) * max(a[ri[f]], 1e-30), tag
            assert np.abs(got["peaks"]["amplitude"] - ref["peaks"]["amplitude"]).max() <= atol, tag
