"""The Node-API shim (napi/pragma_napi.cc) cannot run here (no Node.js in the image); check that it
compiles against the hand-declared Node-API subset, exports the module entry point, and that every
pdsp_* symbol it needs is declared in include/pragma_b200.h and exported by the library."""
import os
import re
import subprocess

from conftest import ROOT


def test_napi_shim_compiles_and_links_against_the_cabi(tmp_path):
    out = tmp_path / "pragma_b200.node"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "napi", "pragma_napi.cc"), "-o", str(out)], check=True)
    syms = subprocess.run(["nm", "-D", str(out)], check=True, capture_output=True, text=True).stdout
    assert re.search(r" T napi_register_module_v1", syms)
    needed = set(re.findall(r" U (pdsp_\w+)", syms))
    assert len(needed) >= 12
    from pragma_dsp_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "pragma_b200.h")).read()
    for s in needed:
        assert hasattr(L, s) and re.search(rf"\b{s}\s*\(", header), s
    napi = set(re.findall(r" U (napi_\w+)", syms))
    declared = set(re.findall(r"\b(napi_\w+)\s*\(", open(os.path.join(ROOT, "napi", "node_api_min.h")).read()))
    assert napi <= declared


def test_ts_host_keeps_the_reference_export_names():
    """SURVEY 8b 'Signatures to keep' - the export names of the four subpaths."""
    want = {
        "ts/core/fft.ts": ["ComplexArray", "createComplexArray", "isPowerOfTwo", "nextPowerOfTwo", "Radix2Fft"],
        "ts/xform/fourier.ts": ["WindowType", "FftSides", "createWindow", "FFT", "magnitude", "phase", "binFrequencies"],
        "ts/public/spectrum.ts": ["SpectrumPeak", "SpectrumResult", "SpectrumOptions", "spectrum"],
        "ts/effect/index.ts": ["FourierService", "Fourier", "FourierLive", "SpectrumFxOptions", "SpectrumFxResult",
                               "spectrumFx", "spectrumStream"],
    }
    for path, names in want.items():
        src = open(os.path.join(ROOT, path)).read()
        for n in names:
            assert re.search(rf"export (const|class|type|interface) {n}\b", src), (path, n)
