"""Mock Node-API host (TEST INFRASTRUCTURE ONLY): executes napi/pragma_napi.cc without Node.js.

mock_napi.cc implements the Node-API subset the addon uses over a toy value model and loads the addon the way Node
does (dlopen + napi_register_module_v1).  `Host` drives it through ctypes: typed arrays are numpy arrays lent to the
addon for the duration of a call, exactly as napi_get_typedarray_info lends JS memory.  The calls the tests make are
the calls ts/*.ts make, argument for argument.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_BUILD = os.path.join(_HERE, "build")
_MOCK = os.path.join(_BUILD, "libmock_napi.so")

# napi_typedarray_type values (stable ABI)
TA_TYPES = {np.dtype(np.uint8): 1, np.dtype(np.int32): 5, np.dtype(np.float32): 7, np.dtype(np.float64): 8}
K_UNDEFINED, K_NULL, K_BOOL, K_NUMBER, K_STRING, K_OBJECT, K_FUNCTION, K_EXTERNAL, K_ARRAYBUFFER, K_TYPEDARRAY = range(10)


class V(int):
    """A napi_value handle of the mock host (distinguishes handles from plain Python numbers)."""


class JsError(Exception):
    """An exception the addon threw with napi_throw_error."""


def _stale(target, srcs):
    return not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(s) for s in srcs)


def build_mock() -> str:
    os.makedirs(_BUILD, exist_ok=True)
    src = os.path.join(_HERE, "mock_napi.cc")
    if _stale(_MOCK, [src, os.path.join(_ROOT, "napi", "node_api_min.h")]):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-Wall", "-Werror", "-fPIC", "-shared", src, "-ldl", "-o", _MOCK],
                       check=True)
    return _MOCK


def build_addon(lib_path: str, tag: str) -> str:
    """pragma_b200.node linked against `lib_path` (libpragma_b200.so, or the emulated C-ABI library on the CPU)."""
    os.makedirs(_BUILD, exist_ok=True)
    out = os.path.join(_BUILD, f"pragma_b200_{tag}.node")
    src = os.path.join(_ROOT, "napi", "pragma_napi.cc")
    deps = [src, os.path.join(_ROOT, "napi", "node_api_min.h"), os.path.join(_ROOT, "include", "pragma_b200.h"), lib_path]
    if _stale(out, deps):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.join(_ROOT, "include"),
                        src, os.path.abspath(lib_path), f"-Wl,-rpath,{os.path.dirname(os.path.abspath(lib_path))}", "-o", out], check=True)
    return out


class Host:
    def __init__(self, lib_path: str, tag: str):
        L = C.CDLL(build_mock(), mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        for name, res, args in (
                ("mock_env_create", vp, []), ("mock_load_addon", vp, [vp, C.c_char_p]), ("mock_number", vp, [vp, C.c_double]),
                ("mock_null", vp, [vp]), ("mock_undefined", vp, [vp]), ("mock_object", vp, [vp]),
                ("mock_object_set", None, [vp, vp, C.c_char_p, vp]), ("mock_typedarray", vp, [vp, C.c_int, vp, C.c_size_t]),
                ("mock_typedarray_on", vp, [vp, vp, C.c_int, C.c_size_t, C.c_size_t]), ("mock_kind", C.c_int, [vp]),
                ("mock_get_number", C.c_double, [vp]), ("mock_buffer_data", vp, [vp]), ("mock_buffer_length", C.c_size_t, [vp]),
                ("mock_has_function", C.c_int, [vp, C.c_char_p]), ("mock_export_names", C.c_int, [vp, C.c_char_p, C.c_size_t]),
                ("mock_call", vp, [vp, vp, C.c_char_p, C.c_size_t, C.POINTER(vp)]), ("mock_error", C.c_char_p, [vp]),
                ("mock_collect", C.c_int, [vp, vp]), ("mock_finalizers_run", C.c_long, [vp]), ("mock_env_destroy", None, [vp, C.c_int])):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        self.L = L
        self.env = L.mock_env_create()
        self.exports = L.mock_load_addon(self.env, build_addon(lib_path, tag).encode())
        if not self.exports:
            raise JsError((L.mock_error(self.env) or b"load failed").decode())
        self._keep = []  # numpy arrays lent to typed-array values stay alive with the host

    # ---- values
    def num(self, x):
        return V(self.L.mock_number(self.env, float(x)))

    def null(self):
        return V(self.L.mock_null(self.env))

    def obj(self, **kw):
        o = V(self.L.mock_object(self.env))
        for k, v in kw.items():
            self.L.mock_object_set(self.env, o, k.encode(), self.num(v))
        return o

    def ta(self, a: np.ndarray):
        """A typed array over `a`'s memory (a must be C-contiguous; its dtype picks the typed-array type)."""
        if a is None:
            return self.null()
        assert a.flags.c_contiguous
        self._keep.append(a)
        return V(self.L.mock_typedarray(self.env, TA_TYPES[a.dtype], a.ctypes.data, a.size))

    def val(self, x):
        if isinstance(x, V):
            return x
        if isinstance(x, (int, float)):
            return self.num(x)
        if isinstance(x, np.ndarray):
            return self.ta(x)
        if x is None:
            return self.null()
        if isinstance(x, dict):
            return self.obj(**x)
        raise TypeError(f"cannot turn {type(x)} into a JS value")

    # ---- calls
    def call(self, name: str, *args):
        argv = (C.c_void_p * max(1, len(args)))(*[self.val(a) for a in args])
        r = self.L.mock_call(self.env, self.exports, name.encode(), len(args), argv)
        if not r:
            raise JsError(self.L.mock_error(self.env).decode())
        return V(r)

    def number(self, v) -> float:
        assert self.L.mock_kind(v) == K_NUMBER
        return self.L.mock_get_number(v)

    def kind(self, v) -> int:
        return self.L.mock_kind(v)

    def export_names(self):
        buf = C.create_string_buffer(4096)
        self.L.mock_export_names(self.exports, buf, 4096)
        return sorted(n for n in buf.value.decode().split("\n") if n)

    def external_float64(self, ab, length, byte_offset=0) -> np.ndarray:
        """numpy view of an (external) ArrayBuffer value - what `new Float64Array(buffer)` sees."""
        n = self.L.mock_buffer_length(ab)
        assert byte_offset + 8 * length <= n
        base = self.L.mock_buffer_data(ab)
        return np.ctypeslib.as_array((C.c_double * length).from_address(base + byte_offset))

    def typedarray_on(self, ab, dtype, length, byte_offset=0):
        return V(self.L.mock_typedarray_on(self.env, ab, TA_TYPES[np.dtype(dtype)], byte_offset, length))

    def collect(self, v) -> bool:
        return bool(self.L.mock_collect(self.env, v))

    def destroy(self, reverse=False):
        self.L.mock_env_destroy(self.env, 1 if reverse else 0)
        self.env = None
