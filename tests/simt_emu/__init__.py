"""Host SIMT emulator harness (TEST ONLY): runs the real kernel source (fft_kernels.cuh) on the CPU.

See emu.cpp.  Used by tests/test_kernel_emulation.py to check the kernels' index math,
synchronisation and epilogue logic against the oracle where no GPU exists.  Not part of the
product: pragma_dsp_b200 never imports it and libpragma_b200.so does not contain it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpdsp_emu.so")
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "pragma_dsp_b200", "csrc")


class R2CParams(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("sample_dtype", C.c_int), ("vec_ok", C.c_int), ("frame_len", C.c_int),
                ("hop", C.c_longlong), ("batch", C.c_longlong), ("window", C.c_void_p), ("tw", C.c_void_p),
                ("post", C.c_void_p), ("out_re", C.c_void_p), ("out_im", C.c_void_p), ("cfull", C.c_int),
                ("amp", C.c_void_p), ("phase", C.c_void_p), ("peaks", C.c_void_p), ("two_sided", C.c_int), ("shift", C.c_int),
                ("scale_edge", C.c_double), ("scale_mid", C.c_double), ("bin_hz", C.c_double),
                ("peer", C.c_void_p * 8), ("n_peers", C.c_int), ("peer_offset", C.c_longlong),
                ("door_flag", C.c_void_p), ("door_count", C.c_void_p), ("door_seq", C.c_uint)]


class C2CParams(C.Structure):
    _fields_ = [("in_re", C.c_void_p), ("in_im", C.c_void_p), ("out_re", C.c_void_p), ("out_im", C.c_void_p),
                ("batch", C.c_longlong), ("tw", C.c_void_p), ("inverse", C.c_int),
                ("door_flag", C.c_void_p), ("door_count", C.c_void_p), ("door_seq", C.c_uint)]


PEAK64 = np.dtype([("index", "<i4"), ("_pad", "<i4"), ("frequency", "<f8"), ("amplitude", "<f8"), ("phase", "<f8")])
PEAK32 = np.dtype([("index", "<i4"), ("frequency", "<f4"), ("amplitude", "<f4"), ("phase", "<f4")])

_lib = None
_CXX = ["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-pthread", "-ffp-contract=off", "-DPDSP_EMU=1", "-I", _HERE] + \
       os.environ.get("PDSP_EMU_EXTRA", "").split()  # e.g. -DPDSP_DERIVE_MID=1 to run an experiment switch under the emulator
_ROOT = os.path.dirname(os.path.dirname(_HERE))


def _stale(target, srcs):
    return not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(s) for s in srcs)


def _headers():
    return [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".h", ".cuh"))] + \
           [os.path.join(_HERE, "cuda_stub.h"), os.path.join(_HERE, "emu_runtime.cc")]


def lib():
    """Kernel-level harness: emu.cpp + emu_runtime.cc -> libpdsp_emu.so."""
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "emu.cpp")] + _headers()
        if _stale(_SO, srcs):
            # emu.cpp is compiled in 8 parts (-DEMU_PART=k: groups of sizes, the tuning variants, the dispatchers)
            # side by side; one translation unit took a minute and a half
            import concurrent.futures as cf
            objdir = os.path.join(_HERE, "build")
            os.makedirs(objdir, exist_ok=True)

            def cc(k):
                obj = os.path.join(objdir, f"emu_part{k}.o")
                if k < 0:
                    obj = os.path.join(objdir, "emu_runtime_k.o")
                    subprocess.run(_CXX + ["-c", os.path.join(_HERE, "emu_runtime.cc"), "-o", obj], check=True)
                else:
                    subprocess.run(_CXX + ["-w", f"-DEMU_PART={k}", "-c", os.path.join(_HERE, "emu.cpp"), "-o", obj], check=True)
                return obj

            with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
                objs = list(ex.map(cc, [-1, 0, 1, 2, 3, 4, 5, 6, 7]))
            subprocess.run(["/usr/bin/g++", "-shared", "-pthread", "-o", _SO] + objs, check=True)
        L = C.CDLL(_SO)
        assert L.emu_params_size(0) == C.sizeof(R2CParams), (L.emu_params_size(0), C.sizeof(R2CParams))
        assert L.emu_params_size(1) == C.sizeof(C2CParams)
        _lib = L
    return _lib


CABI_EMU_SO = os.path.join(_HERE, "libpragma_b200_emu.so")


def build_cabi_emulated():
    """The whole C-ABI library (pragma_b200.cu + a reduced set of inst.cu units) compiled for the host
    against cuda_stub.h, kernels running under the emulator: exercises the host pipeline without a GPU.
    TEST ONLY - loaded by tests/test_host_pipeline_emulated.py through a monkeypatched LIB_PATH."""
    import concurrent.futures as cf
    import re
    srcs = _headers() + [os.path.join(_CSRC, "pragma_b200.cu"), os.path.join(_CSRC, "inst.cu"), os.path.join(_CSRC, "bigfft.cu"), os.path.join(_CSRC, "bigfft2.cu"), os.path.join(_CSRC, "bigfft3.cu"),
                         os.path.join(_ROOT, "include", "pragma_b200.h")]
    if not _stale(CABI_EMU_SO, srcs):
        return CABI_EMU_SO
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    txt = open(os.path.join(_CSRC, "inst_groups.h")).read()
    txt = txt[:txt.index("#else")]
    units = [(os.path.join(_CSRC, "pragma_b200.cu"), os.path.join(objdir, "pragma_b200.o"), []),
             (os.path.join(_CSRC, "bigfft.cu"), os.path.join(objdir, "bigfft.o"), []),
             (os.path.join(_CSRC, "bigfft2.cu"), os.path.join(objdir, "bigfft2.o"), []),
             (os.path.join(_CSRC, "bigfft3.cu"), os.path.join(objdir, "bigfft3.o"), []),
             (os.path.join(_HERE, "emu_runtime.cc"), os.path.join(objdir, "emu_runtime.o"), [])]
    for kind, tag, ctype, lo, hi in re.findall(r"X\((\d), (\w+), (\w+), (\d+), (\d+)\)", txt):
        name = f"launch_{'r2c' if kind == '0' else 'c2c'}_{tag}_{lo}_{hi}"
        units.append((os.path.join(_CSRC, "inst.cu"), os.path.join(objdir, name + ".o"),
                      [f"-DPDSP_INST_KIND={kind}", f"-DPDSP_INST_T={ctype}", f"-DPDSP_INST_LO={lo}",
                       f"-DPDSP_INST_HI={hi}", f"-DPDSP_INST_NAME={name}"]))

    def cc(u):
        src, obj, defs = u
        subprocess.run(_CXX + ["-fvisibility=hidden", "-x", "c++", "-c", src, "-o", obj] + defs, check=True)
        return obj

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, units))
    subprocess.run(["/usr/bin/g++", "-shared", "-pthread", "-o", CABI_EMU_SO] + objs, check=True)
    return CABI_EMU_SO


def _w(k, n):
    """exp(-2*pi*i*k/n) in long double with exact axis/diagonal values (as pragma_b200.cu::twiddle)."""
    k %= n
    if (8 * k) % n == 0:
        o = (8 * k) // n
        s = np.sqrt(np.longdouble(0.5))
        return ([1, s, 0, -s, -1, -s, 0, s][o], [0, -s, -1, -s, 0, s, 1, s][o])
    ang = -2 * np.longdouble(np.pi) * np.longdouble(k) / np.longdouble(n)
    return (np.cos(ang), np.sin(ang))


def pass_table(log2m, rb, dtype):
    """Per-pass twiddles in FftEngine::tw_offset layout (as pragma_b200.cu::upload_pass_twiddles)."""
    rows = []
    npass = 0 if log2m == 0 else -(-log2m // rb)
    for i in range(1, npass):
        bits = rb if i < npass - 1 else log2m - rb * (npass - 1)
        R, Ns = 1 << bits, 1 << (rb * i)
        for s in range(1, R):
            for r in range(Ns):
                rows.append(_w(s * r, Ns * R))
    if not rows:
        rows.append((1, 0))
    return np.ascontiguousarray(np.array(rows, dtype=np.longdouble).astype(dtype))


def tables(n, dtype, variant=0, complex_plan=False):
    """Twiddle and post-pass tables as the product builds them for size n (real plan: M = n/2 points)."""
    m = n if complex_plan else n // 2
    log2m = max(m, 1).bit_length() - 1
    cfg = (C.c_int * 3)()
    rc = lib().emu_cfg(int(dtype == np.float64), log2m, variant, cfg)
    assert rc == 0, f"size {n} not instantiated in the emulator"
    tw = pass_table(log2m, cfg[1], dtype)
    assert cfg[2] == 0 or len(tw) == cfg[2], (len(tw), cfg[2])
    post = np.empty((m // 2 + 1, 2), dtype=dtype)
    if not complex_plan:
        for kk in range(m // 2 + 1):
            r, i = _w(kk, n) if n >= 4 else (np.longdouble(1), np.longdouble(0))
            post[kk] = (i / 2, -r / 2)
    return tw, np.ascontiguousarray(post)


def _p(a):
    return None if a is None else a.ctypes.data


def r2c(samples, n, *, dtype=np.float64, frame_len=None, hop=None, batch=None, window=None, sides="one",
        sample_rate=1.0, raw=False, want=("complex", "amp", "phase", "peak"), cfull=True, nblocks=1,
        specialised=False, variant=0, staged=False):
    """Run r2c_kernel under the emulator. samples: 1-D float32/float64 array."""
    samples = np.ascontiguousarray(samples)
    assert samples.dtype in (np.float32, np.float64)
    frame_len = n if frame_len is None else frame_len
    hop = frame_len if hop is None else hop
    batch = 1 if batch is None else batch
    m = n // 2
    log2m = m.bit_length() - 1
    tw, post = tables(n, dtype, variant)
    win = None if window is None else np.ascontiguousarray(window, dtype=dtype)
    bins = n if sides == "two" else m + 1
    cb = n if cfull else m + 1
    out = {}
    if "complex" in want:
        out["re"] = np.full((batch, cb), np.nan, dtype=dtype)
        out["im"] = np.full((batch, cb), np.nan, dtype=dtype)
    if "amp" in want:
        out["amp"] = np.full((batch, bins), np.nan, dtype=dtype)
    if "phase" in want:
        out["phase"] = np.full((batch, bins), np.nan, dtype=dtype)
    if "peak" in want:
        out["peaks"] = np.zeros(batch, dtype=PEAK64 if dtype == np.float64 else PEAK32)
    p = R2CParams()
    p.samples = _p(samples)
    p.sample_dtype = 1 if samples.dtype == np.float64 else 0
    p.vec_ok = int(samples.ctypes.data % (2 * samples.itemsize) == 0 and hop % 2 == 0)
    p.frame_len, p.hop, p.batch = frame_len, hop, batch
    p.window, p.tw, p.post = _p(win), _p(tw), _p(post)
    p.out_re, p.out_im, p.cfull = _p(out.get("re")), _p(out.get("im")), int(cfull)
    p.amp, p.phase, p.peaks = _p(out.get("amp")), _p(out.get("phase")), _p(out.get("peaks"))
    p.two_sided = int(sides == "two")
    if raw:
        p.scale_edge = p.scale_mid = 1.0
    elif sides == "two":
        p.scale_edge = p.scale_mid = 1.0 / n
    else:
        p.scale_edge, p.scale_mid = 1.0 / n, 2.0 / n
    p.bin_hz = sample_rate / n
    mode = 0
    if specialised:  # compile-time specialised kernel (MD_* bits); preconditions are the caller's job
        assert p.vec_ok and cfull
        mode = (1 if "amp" in want else 0) | (2 if "phase" in want else 0) | (4 if "peak" in want else 0) | \
               (8 if "complex" in want else 0) | (16 if sides == "two" else 0) | (32 if frame_len < n else 0)
        if staged:  # bulk-staged sample loads: 16-byte aligned whole frames
            assert samples.ctypes.data % 16 == 0 and (hop * samples.itemsize) % 16 == 0 and frame_len >= n
            mode |= 64
    if variant:
        assert mode in (1, 5, 4) and n == 1024
        rc = lib().emu_r2c_var(int(dtype == np.float64), variant, C.byref(p), nblocks, mode)
    else:
        rc = lib().emu_r2c(int(dtype == np.float64), log2m, C.byref(p), nblocks, mode)
    assert rc == 0, f"size {n} / mode {mode} not instantiated in the emulator (rc={rc})"
    return out


def c2c(re, im, n, *, dtype=np.float64, inverse=False, nblocks=1):
    re = np.ascontiguousarray(re, dtype=dtype)
    im = None if im is None else np.ascontiguousarray(im, dtype=dtype)
    batch = re.size // n
    tw, _ = tables(n, dtype, complex_plan=True)
    ore = np.full_like(re, np.nan)
    oim = np.full_like(re, np.nan)
    p = C2CParams()
    p.in_re, p.in_im, p.out_re, p.out_im = _p(re), _p(im), _p(ore), _p(oim)
    p.batch, p.tw, p.inverse = batch, _p(tw), int(inverse)
    rc = lib().emu_c2c(int(dtype == np.float64), n.bit_length() - 1, C.byref(p), nblocks)
    assert rc == 0
    return ore, oim
