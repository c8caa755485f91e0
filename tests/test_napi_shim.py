"""The Node-API shim (napi/pragma_napi.cc) EXECUTED under the mock Node-API host of tests/napi_host (this image has
no Node.js): the calls below are the calls ts/*.ts make, argument for argument.  On the CPU the addon is linked against
the emulated C-ABI library (tests/simt_emu: the real host code with kernels under the SIMT emulator); under `-m gpu`
against libpragma_b200.so.  Also: symbol-level checks of the addon and the export names of the TypeScript layer."""
import os
import re
import subprocess

import numpy as np
import pytest

import napi_host
import oracle
from conftest import ROOT

PEAK_F64 = np.dtype([("index", "<i4"), ("_pad", "<i4"), ("frequency", "<f8"), ("amplitude", "<f8"), ("phase", "<f8")])
F32, F64 = 0, 1
WIN = {"rect": 0, "hann": 1, "hamming": 2, "blackman": 3}


def multitone(rng, batch, n):
    t = np.arange(n)
    k = rng.integers(8, n // 2 - 8, size=(batch, 3)) + rng.uniform(-0.25, 0.25, size=(batch, 3))
    a = np.concatenate([np.ones((batch, 1)), rng.uniform(0.1, 0.5, size=(batch, 2))], axis=1)
    ph = rng.uniform(0, 2 * np.pi, size=(batch, 3))
    return (a[:, :, None] * np.sin(2 * np.pi * k[:, :, None] * t[None, None, :] / n + ph[:, :, None])).sum(1)


def test_napi_shim_compiles_and_links_against_the_cabi(tmp_path):
    out = tmp_path / "pragma_b200.node"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "napi", "pragma_napi.cc"), "-o", str(out)], check=True)
    syms = subprocess.run(["nm", "-D", str(out)], check=True, capture_output=True, text=True).stdout
    assert re.search(r" T napi_register_module_v1", syms)
    needed = set(re.findall(r" U (pdsp_\w+)", syms))
    assert len(needed) >= 14 and {"pdsp_apply_window", "pdsp_fft_shift"} <= needed
    from pragma_dsp_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "pragma_b200.h")).read()
    for s in needed:
        assert hasattr(L, s) and re.search(rf"\b{s}\s*\(", header), s
    napi = set(re.findall(r" U (napi_\w+)", syms))
    declared = set(re.findall(r"\b(napi_\w+)\s*\(", open(os.path.join(ROOT, "napi", "node_api_min.h")).read()))
    assert napi <= declared


# every export of the reference's four subpaths (package.json:14-55; src/index.ts, src/core/index.ts,
# src/xform/index.ts, src/xform/fourier.ts, src/public/spectrum.ts, src/effect/index.ts)
REFERENCE_EXPORTS = {
    "ts/core/fft.ts": ["ComplexArray", "createComplexArray", "isPowerOfTwo", "nextPowerOfTwo", "Radix2Fft"],
    "ts/xform/fourier.ts": ["WindowType", "FftSides", "createWindow", "applyWindow", "FFT", "magnitude", "phase", "fftShift",
                            "fftShiftComplex", "binFrequencies"],
    "ts/public/spectrum.ts": ["SpectrumPeak", "SpectrumResult", "SpectrumOptions", "spectrum"],
    "ts/effect/index.ts": ["FourierService", "Fourier", "FourierLive", "SpectrumFxOptions", "SpectrumFxResult",
                           "spectrumFx", "spectrumStream"],
}


def test_ts_host_keeps_the_reference_export_names():
    """SURVEY 8b 'Signatures to keep' - the export names of the four subpaths, the barrels and the subpath map."""
    for path, names in REFERENCE_EXPORTS.items():
        src = open(os.path.join(ROOT, path)).read()
        for n in names:
            assert re.search(rf"export (const|class|type|interface) {n}\b", src), (path, n)
    # barrels as in the reference: src/index.ts:1, src/core/index.ts:3, src/xform/index.ts:1
    assert 'export * from "./public/spectrum.js"' in open(os.path.join(ROOT, "ts", "index.ts")).read()
    assert 'export * from "./fft.js"' in open(os.path.join(ROOT, "ts", "core", "index.ts")).read()
    assert 'export * from "./fourier.js"' in open(os.path.join(ROOT, "ts", "xform", "index.ts")).read()
    import json
    pkg = json.load(open(os.path.join(ROOT, "ts", "package.json")))
    assert set(pkg["exports"]) >= {".", "./core", "./xform", "./xform/fourier", "./effect"}


def test_reference_exports_are_what_the_reference_has():
    """The list above against the reference sources themselves, when the checkout is present (build container only)."""
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present on this machine")
    pairs = {"ts/core/fft.ts": "core/fft.ts", "ts/xform/fourier.ts": "xform/fourier.ts", "ts/public/spectrum.ts": "public/spectrum.ts",
             "ts/effect/index.ts": "effect/index.ts"}
    for mine, theirs in pairs.items():
        src = open(os.path.join(ref, theirs)).read()
        exported = set(re.findall(r"^export (?:const|class|type|interface|function) (\w+)", src, flags=re.M))
        assert exported <= set(REFERENCE_EXPORTS[mine]), (theirs, exported - set(REFERENCE_EXPORTS[mine]))


# ------------------------------------------------------------------------------------------ executed shim
def _exercise(host, *, big=False):
    """The call sequences of ts/core/fft.ts, ts/xform/fourier.ts, ts/public/spectrum.ts and ts/effect/index.ts."""
    assert host.number(host.call("abiVersion")) == 1
    names = host.export_names()
    for want in ("contextCreate", "planGet", "fftForwardReal", "fftForwardComplex", "fftInverse", "magnitude", "phase",
                 "applyWindow", "fftShift", "spectrum", "createWindow", "binFrequencies", "hostAlloc", "ingestOpen", "ingestPush",
                 "ingestFlush", "ingestPop"):
        assert want in names, want
    ctx = host.call("contextCreate", 0)
    assert host.kind(ctx) == napi_host.K_EXTERNAL
    n = 1024
    plan = host.call("planGet", ctx, n, F64)
    rng = np.random.default_rng(5)

    # Radix2Fft.forward(input, out): results land in the caller's arrays (`out` identity is the TS layer's part)
    x = np.sin(np.arange(n, dtype=np.float64))
    re, im = np.full(n, np.nan), np.full(n, np.nan)
    host.call("fftForwardReal", plan, x, re, im)
    rre, rim = oracle.FFT(n).forward(x)
    rel = np.linalg.norm((re - rre.reshape(-1)) + 1j * (im - rim.reshape(-1))) / np.linalg.norm(rre + 1j * rim)
    assert rel <= 1e-11
    # Float32Array input (README.md:11) and a batch of frames in one call
    xb = rng.standard_normal((3, n)).astype(np.float32)
    bre, bim = np.empty((3, n)), np.empty((3, n))
    host.call("fftForwardReal", plan, xb.reshape(-1), bre.reshape(-1), bim.reshape(-1))
    r2, i2 = oracle.FFT(n).forward(xb.astype(np.float64))
    assert np.abs(bre - r2).max() <= 1e-10 and np.abs(bim - i2).max() <= 1e-10
    # forwardComplex -> inverse round trip
    ore, oim, bre2, bim2 = np.empty(n), np.empty(n), np.empty(n), np.empty(n)
    zr, zi = rng.standard_normal(n), rng.standard_normal(n)
    host.call("fftForwardComplex", plan, zr, zi, ore, oim)
    host.call("fftInverse", plan, ore, oim, bre2, bim2)
    assert np.abs(bre2 - zr).max() <= 1e-12 and np.abs(bim2 - zi).max() <= 1e-12

    # xform/fourier: createWindow, applyWindow, magnitude, phase, fftShift, binFrequencies
    w = np.empty(n)
    host.call("createWindow", WIN["hann"], n, w)
    assert (w == oracle.createWindow("hann", n)).all()
    xw = np.empty(n)
    host.call("applyWindow", ctx, x, w, xw)
    assert (xw == x * w).all()
    mag, ph = np.empty(n), np.empty(n)
    host.call("magnitude", ctx, re, im, mag)
    host.call("phase", ctx, re, im, ph)
    assert np.abs(mag - np.hypot(re, im)).max() <= 1e-12 * np.abs(mag).max()
    keep = mag > 1e-6
    d = np.abs(ph - np.arctan2(im, re))
    assert np.minimum(d, np.abs(d - 2 * np.pi))[keep].max() <= 1e-12
    sh = np.empty(n)
    host.call("fftShift", ctx, mag, sh)
    assert (sh == np.fft.fftshift(mag)).all()
    fr = np.empty(n // 2 + 1)
    host.call("binFrequencies", n, 48000.0, 0, fr)
    assert (fr == oracle.binFrequencies(n, 48000.0)).all()

    # spectrum(): one call, amplitude + phase + peak record (ts/public/spectrum.ts)
    frames = multitone(rng, 4, n)
    bins = n // 2 + 1
    amp, phs = np.empty((4, bins)), np.empty((4, bins))
    pk = np.zeros(4 * PEAK_F64.itemsize, dtype=np.uint8)
    host.call("spectrum", plan, frames.reshape(-1),
              dict(frameLen=n, hop=n, batch=4, window=WIN["hann"], sides=0, sampleRate=48000.0), amp.reshape(-1), phs.reshape(-1), pk)
    ref = oracle.spectrum_batch(frames, fftSize=n, sampleRate=48000.0, window="hann")
    assert np.abs(amp - ref["amplitude"]).max() <= 1e-12
    assert (pk.view(PEAK_F64)["index"] == ref["peaks"]["index"]).all()
    assert np.abs(pk.view(PEAK_F64)["frequency"] - ref["peaks"]["frequency"]).max() == 0

    # hostAlloc: pinned ArrayBuffer -> Float64Array views used as `out` (createComplexArray backing store)
    ab = host.call("hostAlloc", ctx, 2 * 8 * n)
    assert host.kind(ab) == napi_host.K_ARRAYBUFFER
    pre = host.typedarray_on(ab, np.float64, n, 0)
    pim = host.typedarray_on(ab, np.float64, n, 8 * n)
    host.call("fftForwardReal", plan, x, pre, pim)
    assert (host.external_float64(ab, n, 0) == re).all() and (host.external_float64(ab, n, 8 * n) == im).all()

    # ingestion ring (ts/effect/index.ts spectrumStream): Float32Array frames in, ordered results out
    f32 = frames.astype(np.float32)
    ring = host.call("ingestOpen", plan, dict(frameLen=n, window=WIN["hann"], sides=0, sampleRate=48000.0, sampleDtype=F32), 1, 0, 1, 2, 3)
    assert host.number(host.call("ingestPush", ring, f32[:3].reshape(-1), 3)) == 3
    host.call("ingestFlush", ring)
    a3 = np.empty((4, bins))
    k3 = np.zeros(4 * PEAK_F64.itemsize, dtype=np.uint8)
    got = int(host.number(host.call("ingestPop", ring, a3.reshape(-1), None, k3, 4)))
    assert got == 3
    ref32 = oracle.spectrum_batch(f32, fftSize=n, sampleRate=48000.0, window="hann")
    assert np.abs(a3[:3] - ref32["amplitude"][:3]).max() <= 1e-12
    assert (k3.view(PEAK_F64)["index"][:3] == ref32["peaks"]["index"][:3]).all()
    return ctx, plan, ring, ab


def _error_mapping(host, ctx, plan, ring):
    """Every bad argument becomes a JS exception (napi_throw_error) before a pointer reaches native code."""
    n = 1024
    x, re, im = np.zeros(n), np.zeros(n), np.zeros(n)
    with pytest.raises(napi_host.JsError, match="lengths do not match"):
        host.call("fftForwardReal", plan, x, np.zeros(n - 1), im)                  # short out.real
    with pytest.raises(napi_host.JsError, match="expected"):
        host.call("fftForwardReal", plan, x, re.astype(np.float32), im)            # Float32Array as out
    with pytest.raises(napi_host.JsError, match="expected a plan handle"):
        host.call("fftForwardReal", ctx, x, re, im)                                # a context where a plan goes
    with pytest.raises(napi_host.JsError, match="expected a plan handle"):
        host.call("fftForwardReal", 7, x, re, im)                                  # a number where a plan goes
    with pytest.raises(napi_host.JsError, match="wrong number of arguments"):
        host.call("fftForwardReal", plan, x)
    with pytest.raises(napi_host.JsError, match="power of two"):                   # library error text, via pdsp_last_error
        host.call("planGet", ctx, 1000, F64)
    with pytest.raises(napi_host.JsError, match="frames exceed"):
        host.call("spectrum", plan, x, dict(frameLen=n, hop=n, batch=2), None, None, None)
    with pytest.raises(napi_host.JsError, match="wrong type or length"):
        host.call("spectrum", plan, x, dict(frameLen=n, hop=n, batch=1), np.zeros(10), None, None)
    with pytest.raises(napi_host.JsError, match="Unsupported window type"):
        host.call("createWindow", 9, 8, np.zeros(8))
    # the ring validates dtype and lengths of what JS hands it
    with pytest.raises(napi_host.JsError, match="sample type"):
        host.call("ingestPush", ring, np.zeros(n, dtype=np.float64), 1)            # ring was opened for Float32Array frames
    with pytest.raises(napi_host.JsError, match="shorter than count"):
        host.call("ingestPush", ring, np.zeros(n, dtype=np.float32), 2)
    with pytest.raises(napi_host.JsError, match="wrong type or length"):
        host.call("ingestPop", ring, np.zeros(513), None, None, 4)                 # room for 1 frame, 4 asked for
    with pytest.raises(napi_host.JsError, match="wrong type or length"):
        host.call("ingestPop", ring, None, None, np.zeros(31, dtype=np.uint8), 1)
    with pytest.raises(napi_host.JsError, match="expected an ingestion ring handle"):
        host.call("ingestFlush", plan)


@pytest.fixture(scope="module")
def emu_lib():
    import simt_emu
    return simt_emu.build_cabi_emulated()


def test_napi_shim_executes_against_the_emulated_cabi(emu_lib):
    host = napi_host.Host(emu_lib, "emu")
    ctx, plan, ring, ab = _exercise(host)
    _error_mapping(host, ctx, plan, ring)
    # ADVICE r1 (use-after-free): the context's finaliser may run BEFORE those of its plans, rings and pinned buffers
    # (GC / teardown order is unspecified) - the boxes keep the context alive until the last of them is gone
    assert host.collect(ctx)
    re, im = np.empty(1024), np.empty(1024)
    host.call("fftForwardReal", plan, np.ones(1024), re, im)   # the plan still works after its context handle was collected
    assert re[0] == 1024.0
    assert host.collect(ring) and host.collect(ab) and host.collect(plan)
    assert not host.collect(plan)                               # finalisers run once
    host.destroy()


@pytest.mark.parametrize("reverse", [False, True])
def test_napi_teardown_order_does_not_matter(emu_lib, reverse):
    host = napi_host.Host(emu_lib, "emu")
    ctx = host.call("contextCreate", 0)
    plans = [host.call("planGet", ctx, n, F64) for n in (64, 1024)]
    ring = host.call("ingestOpen", plans[1], dict(frameLen=1024, window=0, sides=0, sampleRate=1.0, sampleDtype=F32), 1, 1, 1, 4, 2)
    host.call("ingestPush", ring, np.zeros(1024, dtype=np.float32), 1)      # a partial chunk is still pending at teardown
    host.call("hostAlloc", ctx, 4096)
    before = host.L.mock_finalizers_run(host.env)
    env = host.env
    host.destroy(reverse=reverse)                                            # creation order / reverse creation order
    assert before == 0 and env is not None


@pytest.mark.gpu
def test_napi_shim_executes_on_the_gpu():
    from pragma_dsp_b200 import _lib
    host = napi_host.Host(_lib.LIB_PATH, "gpu")
    ctx, plan, ring, ab = _exercise(host)
    _error_mapping(host, ctx, plan, ring)
    host.destroy(reverse=True)
