"""The C-ABI host logic (pragma_b200.cu: plan cache, window tables, chunked staging slots, pinned vs
pageable paths, dispatch generic/specialised, locking) executed WITHOUT a GPU: the library is compiled
for the host against tests/simt_emu/cuda_stub.h and its kernels run under the SIMT emulator.

Test infrastructure only - the product library is not involved; `_lib.LIB_PATH` is monkeypatched for
this module.  The same entry points are tested for real on the B200 by the `-m gpu` tests.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
import simt_emu


@pytest.fixture(scope="module")
def emu_api():
    from pragma_dsp_b200 import _lib
    so = simt_emu.build_cabi_emulated()
    saved = (_lib.LIB_PATH, _lib._lib, _lib._default_ctx)
    _lib.LIB_PATH, _lib._lib, _lib._default_ctx = so, None, None
    try:
        yield _lib
    finally:
        if _lib._default_ctx is not None:
            _lib._default_ctx.close()
        _lib.LIB_PATH, _lib._lib, _lib._default_ctx = saved


@pytest.fixture
def tune(emu_api):
    """Sets tunables of the default context (pdsp_ctx_tune) and restores the defaults afterwards."""
    used = []

    def _set(key, value):
        emu_api.default_context().tune(key, value)
        used.append(key)

    yield _set
    for key in used:
        emu_api.default_context().tune(key, None)


def multitone(rng, batch, n, dtype=np.float64):
    t = np.arange(n)
    k = rng.integers(8, n // 2 - 8, size=(batch, 3)) + rng.uniform(-0.25, 0.25, size=(batch, 3))
    a = np.concatenate([np.ones((batch, 1)), rng.uniform(0.1, 0.5, size=(batch, 2))], axis=1)
    ph = rng.uniform(0, 2 * np.pi, size=(batch, 3))
    x = np.zeros((batch, n))
    for j in range(3):
        x += a[:, j, None] * np.sin(2 * np.pi * k[:, j, None] * t[None, :] / n + ph[:, j, None])
    return x.astype(dtype)


def test_spectrum_host_entry_windowed(emu_api):
    """pdsp_spectrum with a window (regression: plan-cache lock taken inside the pipeline lock)."""
    from pragma_dsp_b200 import spectrum, spectrum_batch
    rng = np.random.default_rng(1)
    x = multitone(rng, 9, 1024)
    for window in ("hann", "rect", "blackman"):
        got = spectrum_batch(x, sampleRate=48000.0, fftSize=1024, window=window)
        ref = oracle.spectrum_batch(x, fftSize=1024, sampleRate=48000.0, window=window)
        assert np.abs(got["amplitude"] - ref["amplitude"]).max() <= 1e-13
        assert (got["peaks"]["index"] == ref["peaks"]["index"]).all()
        d = np.abs(got["phase"] - ref["phase"])
        assert np.minimum(d, np.abs(d - 2 * np.pi))[ref["amplitude"] > 1e-6].max() <= 1e-9
    one = spectrum(x[0], {"sampleRate": 48000.0, "window": "hann"})
    ref1 = oracle.spectrum(x[0], sampleRate=48000.0, window="hann")
    assert one["peak"]["index"] == ref1["peak"]["index"] and len(one["amplitude"]) == 513


def test_chunked_pipeline_many_chunks(emu_api, tune):
    """A job cut into many chunks over the 3 staging slots returns every frame, in order, for pageable
    and pinned buffers; fp32 plan; amplitude + peak outputs."""
    from pragma_dsp_b200 import spectrum_batch
    L = emu_api.lib()
    rng = np.random.default_rng(2)
    # pick_chunk targets ~24 MB per chunk; the chunk_bytes tunable shrinks it so a small job is cut into ~10 chunks
    tune("chunk_bytes", 256 << 10)
    n, batch = 1024, 300  # 4 KB + 4 KB per frame -> 31 frames per chunk
    x = multitone(rng, batch, n, np.float32)
    got = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann", precision="f64", outputs=("amplitude", "peak"))
    ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window="hann", want_phase=False, threads=8)
    assert (got["peaks"]["index"] == ref["peaks"]["index"]).all()
    assert np.abs(got["amplitude"] - ref["amplitude"]).max() <= 1e-12
    # pinned source and destinations: direct copies, identical results
    ctx = emu_api.default_context()
    nbytes = x.nbytes
    hp = C.c_void_p()
    emu_api.check(L.pdsp_host_alloc(ctx.h, nbytes, C.byref(hp)))
    px = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(batch * n,)).reshape(batch, n)
    px[:] = x
    got2 = spectrum_batch(px, sampleRate=48000.0, fftSize=n, window="hann", precision="f64", outputs=("amplitude", "peak"))
    assert (got2["amplitude"] == got["amplitude"]).all() and (got2["peaks"] == got["peaks"]).all()
    emu_api.check(L.pdsp_host_free(ctx.h, hp))


def test_transforms_host_entry(emu_api):
    from pragma_dsp_b200.core import ComplexArray, Radix2Fft, createComplexArray
    from pragma_dsp_b200.xform import FFT, magnitude, phase
    rng = np.random.default_rng(3)
    for n in (1, 2, 8, 64, 1024):
        x = rng.standard_normal(n)
        fft = FFT(n)
        out = createComplexArray(n, 5.0)
        r = fft.forward(x, out)
        assert r is out
        rre, rim = oracle.FFT(n).forward(x)
        assert np.abs(r.real - rre).max() <= 1e-12 and np.abs(r.imag - rim).max() <= 1e-12
        back = fft.inverse(r)
        assert np.abs(back.real - x).max() <= 1e-13 and np.abs(back.imag).max() <= 1e-13
        c = ComplexArray(rng.standard_normal(n), rng.standard_normal(n))
        f = fft.forwardComplex(c)
        fre, fim = oracle.FFT(n).forwardComplex(c.real, c.imag)
        assert np.abs(f.real - fre).max() <= 1e-12 and np.abs(f.imag - fim).max() <= 1e-12
        assert np.abs(magnitude(f) - oracle.magnitude(f.real, f.imag)).max() <= 1e-12
        assert np.abs(phase(f) - oracle.phase(f.real, f.imag)).max() <= 1e-12
    re, im = Radix2Fft(64).forward_batch(rng.standard_normal((300, 64)).astype(np.float32))
    assert re.shape == (300, 64) and np.isfinite(re).all()
    with pytest.raises(ValueError, match="FFT input length 7 != size 8"):
        Radix2Fft(8).forward(np.zeros(7))


def test_generic_fallback_shapes_host_entry(emu_api):
    """Shapes that must take the generic kernel: zero-pad, truncate, odd hop, two-sided, N=1/2, empty batch."""
    from pragma_dsp_b200 import spectrum, spectrum_batch
    rng = np.random.default_rng(4)
    sig = rng.standard_normal(5000).astype(np.float32)
    for n, frame_len, hop, batch, sides in [(64, 40, 7, 50, "one"), (64, 100, 33, 30, "two"), (1024, 1024, 511, 7, "one"),
                                            (2, 2, 2, 9, "one"), (1, 1, 1, 5, "one"), (64, 0, 3, 4, "one")]:
        got = spectrum_batch(sig, fftSize=n, frameLen=frame_len, hop=hop, batch=batch, sampleRate=8000.0, window="hamming", sides=sides)
        ref = oracle.spectrum_batch(sig, fftSize=n, frameLen=frame_len, hop=hop, batch=batch, sampleRate=8000.0, window="hamming", sides=sides)
        assert np.abs(got["amplitude"] - ref["amplitude"]).max() <= 1e-13, (n, frame_len, hop)
        gi, ri = got["peaks"]["index"], ref["peaks"]["index"]
        assert ((gi == ri) | (gi == (n - ri) % n)).all() if sides == "two" else (gi == ri).all()
    r = spectrum(np.array([1.0, 1, 1, 1]), {"sampleRate": 48000, "fftSize": 16})
    assert abs(r["amplitude"][0] - 0.25) < 1e-15 and len(r["amplitude"]) == 9
    r = spectrum([], {"fftSize": 8})  # empty input zero-pads to silence
    assert (r["amplitude"] == 0).all() and r["peak"]["index"] == 0
    empty = spectrum_batch(np.zeros((0, 64)), fftSize=64)
    assert empty["amplitude"].shape == (0, 33)


def test_effect_layer_on_emulated_library(emu_api):
    from pragma_dsp_b200 import spectrum
    from pragma_dsp_b200.effect import FourierLive, spectrumFx, spectrumStream
    rng = np.random.default_rng(5)
    frame = multitone(rng, 1, 1024, np.float32)[0]
    opts = {"sampleRate": 48000.0, "fftSize": 1024, "window": "hann", "sides": "one"}
    with FourierLive() as svc:
        a, b = spectrum(frame, opts), spectrumFx(frame, opts)(svc)
        assert a["peak"] == b["peak"] and (a["amplitude"] == b["amplitude"]).all() and (a["phase"] == b["phase"]).all()
        assert svc.fft(64) is svc.fft(64) and svc.window("hann", 64) is svc.window("hann", 64)
        res = list(spectrumStream([frame, frame[:64], frame], opts | {"fftSize": None}, service=svc, chunk=2))
        assert [len(r["amplitude"]) for r in res] == [513, 33, 513]
        assert res[0]["peak"] == res[2]["peak"] == a["peak"]
        assert list(spectrumStream([], opts, service=svc)) == []


def test_launch_and_plan_accounting(emu_api):
    L = emu_api.lib()
    ctx = emu_api.default_context()
    p1, p2 = ctx.plan(1024, emu_api.F64), ctx.plan(1024, emu_api.F64)
    assert p1.value == p2.value and ctx.plan(1024, emu_api.F32).value != p1.value
    before = ctx.launch_count
    from pragma_dsp_b200 import spectrum_batch
    spectrum_batch(np.zeros((3, 1024)), fftSize=1024)
    assert ctx.launch_count == before + 1
    h = C.c_void_p()
    assert L.pdsp_plan_get(ctx.h, 12, 1, C.byref(h)) != 0 and b"power of two" in L.pdsp_last_error()


@pytest.mark.parametrize("factors,n,env", [("6,6", 4096, {}), ("7,6", 8192, {}), ("6,6,6", 1 << 18, {}),
                                           ("7,7", 1 << 14, {}),
                                           # switchable paths: plain / TMA tile loads, no L2 prefetch, planar work planes
                                           # second generation (TMA tile loads AND stores): planar work planes, one transform per group
                                           ("7,6", 8192, {"big_interleave": "0"}),
                                           ("6,6,6", 1 << 18, {"big_interleave": "0"}),
                                           ("7,6", 8192, {"big_chunk": "1"}),
                                           # three passes, passes 1 + 2 run over groups of 32 k1 blocks (the L2-resident schedule of 2^24)
                                           ("6,6,6", 1 << 18, {"big_resident": "32"}),
                                           # third generation (pipeline passes: split-plane exchange, next tile landing during the
                                           # transform): 256 / 512 / 1024-point passes, two and three passes, mixed with the others
                                           ("8,8", 1 << 16, {"big_pipe": "1"}),
                                           ("10,8", 1 << 18, {"big_pipe": "1"}),
                                           ("8,9", 1 << 17, {"big_pipe": "1"}),
                                           ("6,8,6", 1 << 20, {"big_pipe": "1"}),
                                           ("10,6", 1 << 16, {}),
                                           # middle + last pass fused into one persistent launch (tile-level hand-over through counters)
                                           ("6,6,6", 1 << 18, {"big_fused": "1"}),
                                           ("7,6,6", 1 << 19, {"big_fused": "2"}),
                                           # first generation, its switchable paths: plain / TMA tile loads, no L2 prefetch, planar work planes
                                           ("6,6", 4096, {"big_v2": "0"}),
                                           ("6,6,6", 1 << 18, {"big_v2": "0"}),
                                           ("7,6", 8192, {"big_v2": "0", "big_tma": "0"}),
                                           ("7,6", 8192, {"big_v2": "0", "big_tma": "1"}),
                                           ("6,6", 4096, {"big_v2": "0", "big_interleave": "0", "big_tma": "1"}),
                                           ("7,6", 8192, {"big_v2": "0", "big_interleave": "0", "big_tma": "0",
                                                          "big_prefetch": "0"})])
def test_multipass_large_fft_emulated(emu_api, tune, factors, n, env):
    """K2: the multi-pass (four-step / six-step) path, forced onto small sizes with the big_factors tunable so the
    emulator can run it: forward, inverse, real-input forward, batch of 2."""
    from pragma_dsp_b200.core import Radix2Fft
    for k, v in env.items():
        tune(k, v)
    if not (n > 8192 and factors == "7,7"):  # 2^14 = 128 x 128 is the default split
        tune("big_factors", factors)
    rng = np.random.default_rng(n)
    batch = 2 if n <= 8192 else 1
    re, im = rng.standard_normal((batch, n)), rng.standard_normal((batch, n))
    fft = Radix2Fft(n)
    ore, oim = fft.complex_batch(re, im)
    ref = np.fft.fft(re + 1j * im, axis=1)
    tol = 1e-12 * np.log2(n)
    for f in range(batch):
        assert np.linalg.norm((ore[f] + 1j * oim[f]) - ref[f]) / np.linalg.norm(ref[f]) <= tol
    if n >= 1 << 20:  # the emulator runs a million-point transform in ~15 s: forward only (the smaller cases cover the rest)
        return
    bre, bim = fft.complex_batch(ore, oim, inverse=True)
    assert np.abs(bre - re).max() <= 1e-12 and np.abs(bim - im).max() <= 1e-12
    if n <= 8192:
        rre, rim = oracle.FFT(n).forwardComplex(re, im)
        assert np.linalg.norm((ore + 1j * oim) - (rre + 1j * rim)) / np.linalg.norm(rre + 1j * rim) <= tol
    if n > 16384 or factors == "6,6":
        xr = rng.standard_normal(n)
        out = fft.forward(xr) if n > 16384 else None
        if out is not None:
            refr = np.fft.fft(xr)
            assert np.linalg.norm((out.real + 1j * out.imag) - refr) / np.linalg.norm(refr) <= tol


@pytest.mark.parametrize("sides,frame_len", [("one", 4096), ("two", 4096), ("one", 3000)])
def test_large_spectrum_path_emulated(emu_api, tune, sides, frame_len):
    """spectrum() beyond one CTA (N > 16384 on the device): window -> multi-pass transform -> epilogue kernel with
    findPeak.  Forced onto N=4096 with the big_factors tunable so the emulator can run it; compared with the oracle."""
    from pragma_dsp_b200 import spectrum_batch
    tune("big_factors", "6,6")
    rng = np.random.default_rng(11)
    n, batch = 4096, 3
    x = multitone(rng, batch, frame_len)
    x[2] = 0.0  # a silent frame: findPeak falls back to the DC bin
    for window in ("hann", "rect"):
        got = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window=window, sides=sides)
        ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window=window, sides=sides)
        assert got["amplitude"].shape == ref["amplitude"].shape
        assert np.abs(got["amplitude"] - ref["amplitude"]).max() <= 1e-12
        if sides == "one":  # two-sided mirrored bins tie to the last bit: the index may be k or N-k
            assert (got["peaks"]["index"] == ref["peaks"]["index"]).all()
        else:
            gi, ri = got["peaks"]["index"], ref["peaks"]["index"]
            assert ((gi == ri) | (gi == n - ri)).all()
        assert np.abs(got["peaks"]["amplitude"] - ref["peaks"]["amplitude"]).max() <= 1e-12
        d = np.abs(got["phase"] - ref["phase"])
        assert np.minimum(d, np.abs(d - 2 * np.pi))[ref["amplitude"] > 1e-6].max() <= 1e-8
    # peaks only, fp32 plan
    got = spectrum_batch(x.astype(np.float32), sampleRate=48000.0, fftSize=n, window="hann", sides=sides, precision="f32",
                         outputs=("peak",))
    ref = oracle.spectrum_batch(x.astype(np.float32), fftSize=n, sampleRate=48000.0, window="hann", sides=sides)
    if sides == "one":
        assert (got["peaks"]["index"] == ref["peaks"]["index"]).all()
    assert np.abs(got["peaks"]["amplitude"] - ref["peaks"]["amplitude"]).max() <= 2e-6


def test_fused_fftshift_two_sided_emulated(emu_api, tune):
    """fftShift fused into the two-sided stores (desc.fft_shift) equals fftShift() applied afterwards,
    in the fused single-CTA kernel and in the large-N epilogue; one-sided + shift is rejected."""
    from pragma_dsp_b200 import spectrum_batch
    from pragma_dsp_b200.xform import fftShift
    rng = np.random.default_rng(12)
    for n in (2, 8, 64, 1024):  # log2(N/2) = 0, 2, 5, 9: sizes the emulated library instantiates
        x = multitone(rng, 3, n) if n >= 64 else rng.standard_normal((3, n))
        plain = spectrum_batch(x, sampleRate=8000.0, fftSize=n, window="hann", sides="two")
        fused = spectrum_batch(x, sampleRate=8000.0, fftSize=n, window="hann", sides="two", shift=True)
        for f in range(3):
            assert np.array_equal(fused["amplitude"][f], fftShift(plain["amplitude"][f])), n
            assert np.array_equal(fused["phase"][f], fftShift(plain["phase"][f])), n
        assert (fused["peaks"] == plain["peaks"]).all()
    with pytest.raises(Exception, match="two-sided"):
        spectrum_batch(multitone(rng, 1, 64), fftSize=64, sides="one", shift=True)
    tune("big_factors", "6,6")
    x = multitone(rng, 2, 4096)
    plain = spectrum_batch(x, sampleRate=8000.0, fftSize=4096, sides="two")
    fused = spectrum_batch(x, sampleRate=8000.0, fftSize=4096, sides="two", shift=True)
    assert np.array_equal(fused["amplitude"], np.fft.fftshift(plain["amplitude"], axes=1))
    assert np.array_equal(fused["phase"], np.fft.fftshift(plain["phase"], axes=1))


def test_ingestion_ring_emulated(emu_api):
    """frames pushed one at a time / in blocks come back in order, bit-identical to spectrum_batch;
    ring-full back-pressure, partial-chunk flush, pop smaller than a chunk, and spectrumStream on top of the ring."""
    from pragma_dsp_b200 import IngestRing, spectrum_batch
    from pragma_dsp_b200.effect import FourierLive, spectrumStream
    rng = np.random.default_rng(21)
    n, total = 1024, 23
    x = multitone(rng, total, n, np.float32)
    ref = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann")
    with IngestRing(n, sampleRate=48000.0, window="hann", sample_dtype=np.float32, framesPerChunk=4, depth=3) as ring:
        amp, ph, pk = [], [], []
        pushed = 0
        while pushed < total:
            k = ring.push(x[pushed:pushed + 5])  # blocks of 5 against chunks of 4: chunks fill across pushes
            pushed += k
            if k == 0:  # ring full (3 chunks of 4 in flight): collect 3 frames - less than a chunk - and go on
                out = ring.pop(3)
                assert out["count"] == 3
                amp.append(out["amplitude"]), ph.append(out["phase"]), pk.append(out["peaks"])
        ring.flush()
        while True:
            out = ring.pop(7)
            if out["count"] == 0:
                break
            amp.append(out["amplitude"]), ph.append(out["phase"]), pk.append(out["peaks"])
        amp, ph, pk = np.concatenate(amp), np.concatenate(ph), np.concatenate(pk)
    assert amp.shape == ref["amplitude"].shape
    assert np.array_equal(amp, ref["amplitude"]) and np.array_equal(ph, ref["phase"]) and (pk == ref["peaks"]).all()
    # the stream operator: ordered, 1:1, frames of another length in the middle, an empty stream
    svc = FourierLive()
    frames = [x[i] for i in range(9)] + [x[9][:64]] + [x[i] for i in range(10, 14)]
    outs = list(spectrumStream(frames, {"sampleRate": 48000.0, "window": "hann"}, service=svc, chunk=4))
    assert len(outs) == len(frames)
    for i, o in enumerate(outs):
        r = oracle.spectrum(frames[i], sampleRate=48000.0, window="hann")
        assert o["peak"]["index"] == r["peak"]["index"], i
        assert np.abs(o["amplitude"] - r["amplitude"]).max() <= 1e-12
    assert list(spectrumStream([], {}, service=svc)) == []
    with pytest.raises(Exception, match="depth"):
        IngestRing(n, depth=1)


def test_ring_pinned_push_ready_and_threaded_copies_emulated(emu_api, tune):
    """Round-2 ring features: pdsp_ingest_push_pinned (no host copy: the H2D count stays flat while frames are taken),
    pdsp_ingest_ready (non-blocking progress), pop_into caller-owned arrays, the host copy pool on large pushes, and a
    chunk refusing to mix copied and pinned frames."""
    from pragma_dsp_b200 import IngestRing, spectrum_batch
    L = emu_api.lib()
    ctx = emu_api.default_context()
    rng = np.random.default_rng(31)
    n, total = 1024, 96
    x = multitone(rng, total, n, np.float32)
    ref = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann", outputs=("amplitude", "peak"))
    # pinned source: allocate through the library, view as numpy
    hp = C.c_void_p()
    assert L.pdsp_host_alloc(ctx.h, x.nbytes, C.byref(hp)) == 0
    xp = np.ctypeslib.as_array((C.c_float * x.size).from_address(hp.value)).reshape(x.shape)
    xp[:] = x
    amp = np.empty((total, n // 2 + 1))
    pk = np.zeros(total, dtype=emu_api.PEAK_F64)
    with IngestRing(n, sampleRate=48000.0, window="hann", sample_dtype=np.float32, outputs=("amplitude", "peak"),
                    framesPerChunk=32, depth=3) as ring:
        assert ring.ready() == (0, 0, 0)
        assert ring.push_pinned(xp[:40]) == 40          # one full chunk sent, 8 frames waiting
        fin, fly, pend = ring.ready()
        assert fin + fly == 32 and pend == 8
        with pytest.raises(Exception, match="mix"):
            ring.push(x[40:41])                          # a copied frame into the chunk that holds pinned ones
        assert ring.push_pinned(xp[40:]) == 56
        ring.flush()
        got = 0
        while got < total:
            k = ring.pop_into(amp[got:], None, pk[got:], max_frames=min(50, total - got))
            assert k > 0
            got += k
        assert ring.ready() == (0, 0, 0)
        with pytest.raises(ValueError):
            ring.pop_into(amp.astype(np.float32))
        with pytest.raises(Exception, match="pinned"):
            ring.push_pinned(x[:1])                      # pageable memory is refused, not silently copied
    assert np.array_equal(amp, ref["amplitude"]) and (pk == ref["peaks"]).all()
    L.pdsp_host_free(ctx.h, hp)
    # threaded host copies: pushes of 1 MB and more are split over the pool's threads (copy_threads tunable)
    for threads in (0, 3):
        tune("copy_threads", threads)
        with IngestRing(n, sampleRate=48000.0, window="hann", sample_dtype=np.float32, outputs=("amplitude", "peak"),
                        framesPerChunk=96, depth=2) as ring:
            assert ring.push(np.tile(x, (1, 1))) == total and ring.ready()[0] + ring.ready()[1] == total
            out = ring.pop(total)
        assert out["count"] == total and np.array_equal(out["amplitude"], ref["amplitude"])


def test_spectrum_stream_latency_contract_emulated(emu_api):
    """spectrumStream must not sit on results.  With an idle ring every frame is sent as soon as it arrives,
    so a lazy source sees result i before it is asked for frame i + 2 at the latest; lockstep=True gives the reference's
    strict one-in-one-out order (a source that depends on the previous result does not deadlock)."""
    from pragma_dsp_b200.effect import FourierLive, spectrumStream
    rng = np.random.default_rng(41)
    n = 1024  # a size the emulated library instantiates
    x = multitone(rng, 12, n, np.float32)
    svc = FourierLive()
    produced, consumed = [], []

    def source():
        for i in range(12):
            produced.append(i)
            yield x[i]

    for r in spectrumStream(source(), {"sampleRate": 8000.0}, service=svc, chunk=64):
        consumed.append(len(produced))                  # frames the source had produced when result len(consumed)-1 came out
    assert len(consumed) == 12
    assert all(c - i <= 2 for i, c in enumerate(consumed)), consumed   # never `chunk` frames behind
    # feedback source: frame i+1 is only produced after result i was seen
    seen = []

    def feedback():
        for i in range(6):
            assert len(seen) == i, "the stream asked for frame %d before yielding result %d" % (i, i - 1)
            yield x[i]

    for r in spectrumStream(feedback(), {"sampleRate": 8000.0}, service=svc, lockstep=True):
        seen.append(r["peak"]["index"])
    ref = [oracle.spectrum(x[i], sampleRate=8000.0)["peak"]["index"] for i in range(6)]
    assert seen == ref


def test_single_call_fast_lane_emulated(emu_api, tune):
    """one-frame calls (Radix2Fft.forward, spectrum()) take the fast lane - one launch through the
    host-mapped buffer, no staging slots - and give the same bits as the pipeline; jobs beyond 256 KB do not."""
    from pragma_dsp_b200 import spectrum_batch
    from pragma_dsp_b200.core import ComplexArray, Radix2Fft
    ctx = emu_api.default_context()
    rng = np.random.default_rng(51)
    n = 1024
    x = multitone(rng, 3, n)
    f0, h0 = ctx.fast_call_count, emu_api.lib().pdsp_stub_counter(0)
    fft = Radix2Fft(n)
    a = fft.forward(x[0])
    b = fft.inverse(fft.forwardComplex(ComplexArray(x[1], x[2])))
    s = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann")
    assert ctx.fast_call_count == f0 + 4 and emu_api.lib().pdsp_stub_counter(0) == h0   # no H2D copy calls at all
    assert np.abs(b.real - x[1]).max() <= 1e-12 and np.abs(b.imag - x[2]).max() <= 1e-12
    tune("fast", 0)
    a2 = fft.forward(x[0])
    s2 = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann")
    assert ctx.fast_call_count == f0 + 4
    assert np.array_equal(a.real, a2.real) and np.array_equal(a.imag, a2.imag)
    assert np.array_equal(s["amplitude"], s2["amplitude"]) and np.array_equal(s["phase"], s2["phase"]) and (s["peaks"] == s2["peaks"]).all()
    tune("fast", 1)
    big = multitone(rng, 40, n)                           # 40 x (8 KB in + 8 KB out) > 256 KB: the staging pipeline
    f1 = ctx.fast_call_count
    spectrum_batch(big, sampleRate=48000.0, fftSize=n)
    assert ctx.fast_call_count == f1


def test_device_group_emulated(emu_api, monkeypatch):
    """pdsp_group_* with two pretend devices (PDSP_STUB_DEVICES: both are this host's memory): the host-sharded form
    returns exactly what one device returns; the device-resident form leaves ALL peak records in EVERY device's gather
    buffer (fused peer stores) and the root's row buffers hold every block's amplitude / phase rows."""
    from pragma_dsp_b200 import DeviceGroup, spectrum_batch
    from pragma_dsp_b200._lib import F64, PEAK_F64, SIDES, WINDOWS, SpectrumDesc
    from pragma_dsp_b200.group import block_of
    monkeypatch.setenv("PDSP_STUB_DEVICES", "2")
    rng = np.random.default_rng(61)
    n, batch = 1024, 11  # blocks of 6 and 5 frames
    x = multitone(rng, batch, n)
    ref = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann")
    with DeviceGroup([0, 1]) as grp:
        got = grp.spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann")
        for key in ("amplitude", "phase"):
            assert np.array_equal(got[key], ref[key]), key
        assert (got["peaks"] == ref["peaks"]).all()
        # STFT addressing across the block boundary (hop < N: the second block starts inside the first block's samples)
        s = multitone(rng, 1, 10 * 256 + n)[0]
        a = grp.spectrum_batch(s, sampleRate=48000.0, fftSize=n, window="hann", frameLen=n, hop=256, batch=11, outputs=("amplitude",))
        b = spectrum_batch(s, sampleRate=48000.0, fftSize=n, window="hann", frameLen=n, hop=256, batch=11, outputs=("amplitude",))
        assert np.array_equal(a["amplitude"], b["amplitude"])
        # device-resident: "device" memory is host memory under the stub
        bins = n // 2 + 1
        blocks = [block_of(batch, 2, i) for i in range(2)]
        d_x = [np.ascontiguousarray(x[f0:f0 + nf]) for f0, nf in blocks]
        d_amp = [np.empty((nf, bins)) for _, nf in blocks]
        d_ph = [np.empty((nf, bins)) for _, nf in blocks]
        d_pk = [np.zeros(batch, dtype=PEAK_F64) for _ in blocks]
        amp_all, ph_all = np.empty((batch, bins)), np.empty((batch, bins))
        desc = SpectrumDesc(sample_dtype=F64, frame_len=n, hop=n, batch=batch, window=WINDOWS["hann"], sides=SIDES["one"],
                            sample_rate=48000.0, raw_magnitude=0, fft_shift=0)
        ptr = lambda arrs: [a_.ctypes.data for a_ in arrs]  # noqa: E731
        grp.spectrum_dev(n, F64, desc, ptr(d_x), ptr(d_amp), ptr(d_ph), ptr(d_pk), gather_root=1,
                         d_amplitude_all=amp_all.ctypes.data, d_phase_all=ph_all.ctypes.data)
        grp.sync()
        for i in range(2):
            assert (d_pk[i] == ref["peaks"]).all(), i            # every device holds every frame's record
        assert np.array_equal(amp_all, ref["amplitude"]) and np.array_equal(ph_all, ref["phase"])
        with pytest.raises(Exception, match="out of range"):
            grp.spectrum_dev(n, F64, desc, ptr(d_x), None, None, ptr(d_pk), gather_root=2)
    with pytest.raises(Exception, match="listed twice"):
        DeviceGroup([0, 0])


def test_fused_peer_scatter_emulated(emu_api):
    """pdsp_spectrum_dev_gather on the emulated library: two 'peer' buffers receive every record."""
    from pragma_dsp_b200._lib import F64, PEAK_F64, SIDES, WINDOWS, SpectrumDesc
    L = emu_api.lib()
    ctx = emu_api.default_context()
    rng = np.random.default_rng(6)
    n, batch, offset = 1024, 9, 3
    x = multitone(rng, batch, n)
    local = np.zeros(batch, dtype=PEAK_F64)
    tgt = [np.zeros(offset + batch + 2, dtype=PEAK_F64) for _ in range(2)]
    peers = (C.c_void_p * 8)()
    peers[0], peers[1] = tgt[0].ctypes.data, tgt[1].ctypes.data
    d = SpectrumDesc(sample_dtype=F64, frame_len=n, hop=n, batch=batch, window=WINDOWS["hann"], sides=SIDES["one"],
                     sample_rate=48000.0, raw_magnitude=0)
    emu_api.check(L.pdsp_spectrum_dev_gather(ctx.plan(n, F64), C.byref(d), C.c_void_p(x.ctypes.data), None, None,
                                             C.c_void_p(local.ctypes.data), peers, 2, offset, None))
    ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=48000.0, window="hann")
    assert (local["index"] == ref["peaks"]["index"]).all()
    assert np.abs(local["amplitude"] - ref["peaks"]["amplitude"]).max() <= 1e-13
    assert np.abs(local["phase"] - ref["peaks"]["phase"]).max() <= 1e-9
    for t_ in tgt:
        assert (t_[offset:offset + batch] == local).all() and (t_[:offset]["index"] == 0).all()


def test_apply_window_fftshift_stft_emulated(emu_api):
    from pragma_dsp_b200 import stft
    from pragma_dsp_b200.core import ComplexArray
    from pragma_dsp_b200.xform import applyWindow, createWindow, fftShift, fftShiftComplex
    rng = np.random.default_rng(8)
    x = rng.standard_normal(1024)
    w = createWindow("hann", 1024)
    assert (applyWindow(x, w) == x * w).all()  # src/xform/fourier.ts:54-67
    with pytest.raises(ValueError, match="Window length must match input length."):
        applyWindow(x, w[:10])
    for n in (1, 2, 7, 8):  # src/xform/fourier.ts:122-134, odd and even
        v = np.arange(n, dtype=np.float64)
        assert (fftShift(v) == oracle.fftShift(v)).all()
    c = fftShiftComplex(ComplexArray(np.arange(8.0), -np.arange(8.0)))
    assert (c.real == np.roll(np.arange(8.0), -4)).all() and (c.imag == -c.real).all()
    sig = rng.standard_normal(1024 + 7 * 256).astype(np.float32)
    r = stft(sig, fftSize=1024, hopSize=256, window="hann", sampleRate=48000.0, outputs=("amplitude", "peak"))
    ref = oracle.spectrum_batch(sig, fftSize=1024, frameLen=1024, hop=256, batch=8, sampleRate=48000.0, window="hann")
    assert r["amplitude"].shape == (8, 513) and np.abs(r["amplitude"] - ref["amplitude"]).max() <= 1e-13
    assert (r["peaks"]["index"] == ref["peaks"]["index"]).all() and r["times"][1] == 256 / 48000.0
    assert stft(sig[:100], fftSize=1024, hopSize=256)["amplitude"].shape == (0, 513)


def test_randomised_shapes_against_oracle(emu_api):
    """Seeded random walk over the spectrum() parameter space through the whole emulated C-ABI path (specialised and
    generic kernels, zero-padding, truncation, overlapping / gapped / odd hops, both sides, every window, output
    subsets, fused shift): amplitude, peak record and phase against the oracle."""
    from pragma_dsp_b200 import spectrum_batch
    rng = np.random.default_rng(2024)
    sizes = [1, 2, 4, 8, 16, 32, 64, 1024]  # log2(N/2) in the emulated library's instantiation list
    for case in range(40):
        n = int(rng.choice(sizes))
        batch = int(rng.integers(1, 6))
        frame_len = int(rng.choice([n, max(1, n - 1), max(1, n // 2 + 1), n + 3, 1]))
        hop = int(rng.choice([frame_len, max(1, frame_len // 2), frame_len + 5, max(1, frame_len - 1)]))
        window = str(rng.choice(["rect", "hann", "hamming", "blackman"]))
        sides = str(rng.choice(["one", "two"]))
        dtype = np.float64 if rng.integers(0, 2) else np.float32
        outputs = [("amplitude", "phase", "peak"), ("amplitude",), ("peak",), ("amplitude", "peak"), ("phase",)][int(rng.integers(0, 5))]
        shift = bool(sides == "two" and rng.integers(0, 2))
        total = (batch - 1) * hop + frame_len
        x = rng.standard_normal(total).astype(dtype)
        if n >= 16:
            k0 = int(rng.integers(1, n // 2))
            x += (3.0 * np.sin(2 * np.pi * k0 * np.arange(total) / n)).astype(dtype)
        tag = (case, n, batch, frame_len, hop, window, sides, dtype.__name__, outputs, shift)
        got = spectrum_batch(x, sampleRate=8000.0, fftSize=n, window=window, sides=sides, frameLen=frame_len, hop=hop,
                             batch=batch, outputs=outputs, shift=shift)
        ref = oracle.spectrum_batch(x, fftSize=n, sampleRate=8000.0, window=window, sides=sides, frameLen=frame_len,
                                    hop=hop, batch=batch)
        ramp, rph = ref["amplitude"], ref["phase"]
        if shift:
            ramp, rph = np.fft.fftshift(ramp, axes=1), np.fft.fftshift(rph, axes=1)
        if "amplitude" in outputs:
            assert np.abs(got["amplitude"] - ramp).max() <= 1e-12 * max(1.0, np.abs(ramp).max()), tag
        if "phase" in outputs:
            strong = ramp > 1e-6 * max(1.0, ramp.max())
            d = np.abs(got["phase"] - rph)
            if strong.any():
                assert np.minimum(d, np.abs(d - 2 * np.pi))[strong].max() <= 1e-8, tag
        if "peak" in outputs:
            gi, ri = got["peaks"]["index"], ref["peaks"]["index"]
            if sides == "one":
                # near-ties between neighbouring bins of pure noise can flip on the last ulp: accept them only there
                bad = gi != ri
                for f in np.nonzero(bad)[0]:
                    a = ref["amplitude"][f]
                    assert abs(a[gi[f]] - a[ri[f]]) <= 1e-12 * a[ri[f]], tag
            else:
                assert ((gi == ri) | (gi == (n - ri) % n)).all() or n < 4, tag
            assert np.abs(got["peaks"]["amplitude"] - ref["peaks"]["amplitude"]).max() <= 1e-12 * max(1.0, ramp.max()), tag
