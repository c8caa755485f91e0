"""ctypes binding of libpragma_b200.so (include/pragma_b200.h).

The shared library IS the product: if it is missing, or no B200 is visible, importing or using
the package fails loudly.  There is no CPU fallback and nothing here imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpragma_b200.so")

F32, F64 = 0, 1
WINDOWS = {"rect": 0, "hann": 1, "hamming": 2, "blackman": 3}
SIDES = {"one": 0, "two": 1}

PEAK_F64 = np.dtype([("index", "<i4"), ("_pad", "<i4"), ("frequency", "<f8"), ("amplitude", "<f8"), ("phase", "<f8")])
PEAK_F32 = np.dtype([("index", "<i4"), ("frequency", "<f4"), ("amplitude", "<f4"), ("phase", "<f4")])


class SpectrumDesc(C.Structure):
    _fields_ = [("sample_dtype", C.c_int32), ("frame_len", C.c_int32), ("hop", C.c_int64), ("batch", C.c_int64),
                ("window", C.c_int32), ("sides", C.c_int32), ("sample_rate", C.c_double),
                ("raw_magnitude", C.c_int32), ("fft_shift", C.c_int32)]


class PragmaB200Error(RuntimeError):
    pass


# every symbol include/pragma_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
SYMBOLS = {
    "pdsp_abi_version": (C.c_int, []),
    "pdsp_last_error": (C.c_char_p, []),
    "pdsp_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pdsp_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "pdsp_ctx_destroy": (C.c_int, [_vp]),
    "pdsp_ctx_sync": (C.c_int, [_vp]),
    "pdsp_ctx_device": (C.c_int, [_vp]),
    "pdsp_ctx_tune": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "pdsp_plan_release_stream": (C.c_int, [_vp, _vp]),
    "pdsp_ctx_sm_count": (C.c_int, [_vp]),
    "pdsp_ctx_launch_count": (_i64, [_vp]),
    "pdsp_ctx_fast_call_count": (_i64, [_vp]),
    "pdsp_ctx_ping": (C.c_int, [_vp]),
    "pdsp_is_power_of_two": (C.c_int, [_i32]),
    "pdsp_next_power_of_two": (_i32, [_i32]),
    "pdsp_create_window": (C.c_int, [C.c_int, _i32, _dp]),
    "pdsp_bin_frequencies": (C.c_int, [_i32, C.c_double, C.c_int, _dp, C.POINTER(_i32)]),
    "pdsp_plan_get": (C.c_int, [_vp, _i32, C.c_int, C.POINTER(_vp)]),
    "pdsp_plan_size": (_i32, [_vp]),
    "pdsp_plan_precision": (C.c_int, [_vp]),
    "pdsp_fft_forward_real": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp, _vp]),
    "pdsp_fft_forward_complex": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "pdsp_fft_inverse": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "pdsp_magnitude": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "pdsp_phase": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "pdsp_apply_window": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "pdsp_fft_shift": (C.c_int, [_vp, _vp, _i64, _vp]),
    "pdsp_spectrum": (C.c_int, [_vp, C.POINTER(SpectrumDesc), _vp, _vp, _vp, _vp]),
    "pdsp_spectrum_dev": (C.c_int, [_vp, C.POINTER(SpectrumDesc), _vp, _vp, _vp, _vp, _vp]),
    "pdsp_spectrum_dev_gather": (C.c_int, [_vp, C.POINTER(SpectrumDesc), _vp, _vp, _vp, _vp, C.POINTER(_vp), C.c_int, _i64, _vp]),
    "pdsp_ipc_export": (C.c_int, [_vp, _vp, C.c_char_p]),
    "pdsp_ipc_open": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "pdsp_ipc_close": (C.c_int, [_vp, _vp]),
    "pdsp_fft_forward_real_dev": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp, _vp, C.c_int, _vp]),
    "pdsp_fft_complex_dev": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, C.c_int, _vp]),
    "pdsp_complex_mul_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_double, _i64, _vp, _vp, _vp]),
    "pdsp_dev_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "pdsp_dev_free": (C.c_int, [_vp, _vp]),
    "pdsp_host_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "pdsp_host_free": (C.c_int, [_vp, _vp]),
    "pdsp_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "pdsp_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "pdsp_ingest_open": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _i64, C.c_int, C.POINTER(C.c_void_p)]),
    "pdsp_ingest_push": (C.c_int, [_vp, _vp, _i64, _i64, C.POINTER(_i64)]),
    "pdsp_ingest_push_pinned": (C.c_int, [_vp, _vp, _i64, _i64, C.POINTER(_i64)]),
    "pdsp_ingest_ready": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "pdsp_ingest_flush": (C.c_int, [_vp]),
    "pdsp_ingest_pop": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    "pdsp_ingest_close": (C.c_int, [_vp]),
    "pdsp_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "pdsp_group_destroy": (C.c_int, [_vp]),
    "pdsp_group_size": (C.c_int, [_vp]),
    "pdsp_group_ctx": (_vp, [_vp, C.c_int]),
    "pdsp_group_spectrum": (C.c_int, [_vp, _i32, C.c_int, C.POINTER(SpectrumDesc), _vp, _vp, _vp, _vp]),
    "pdsp_group_spectrum_dev": (C.c_int, [_vp, _i32, C.c_int, C.POINTER(SpectrumDesc), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                          C.POINTER(_vp), C.c_int, _vp, _vp]),
    "pdsp_group_sync": (C.c_int, [_vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA extension; raise (never fall back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PragmaB200Error(
                f"{LIB_PATH} is missing: build the CUDA extension with `python -m pragma_dsp_b200.build` "
                "(pragma_dsp_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise PragmaB200Error(lib().pdsp_last_error().decode())


class Context:
    """pdsp_ctx: one per device."""

    def __init__(self, device: int | None = None):
        if device is None:
            device = int(os.environ.get("PDSP_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        h = C.c_void_p()
        check(lib().pdsp_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        self.sm_count = lib().pdsp_ctx_sm_count(h)

    def plan(self, size: int, precision: int = F64) -> C.c_void_p:
        p = C.c_void_p()
        check(lib().pdsp_plan_get(self.h, int(size), int(precision), C.byref(p)))
        return p

    def sync(self) -> None:
        check(lib().pdsp_ctx_sync(self.h))

    def tune(self, key: str, value=None) -> None:
        """pdsp_ctx_tune: set one tunable (tests and sweeps); None restores the default."""
        check(lib().pdsp_ctx_tune(self.h, key.encode(), None if value is None else str(value).encode()))

    @property
    def launch_count(self) -> int:
        return int(lib().pdsp_ctx_launch_count(self.h))

    @property
    def fast_call_count(self) -> int:
        return int(lib().pdsp_ctx_fast_call_count(self.h))

    def close(self) -> None:
        if self.h:
            lib().pdsp_ctx_destroy(self.h)
            self.h = None


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def ptr(a) -> C.c_void_p:
    return None if a is None else C.c_void_p(a.ctypes.data)
