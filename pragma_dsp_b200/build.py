"""Builds libpragma_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile).

    python -m pragma_dsp_b200.build [--force] [--jobs N]

Every kernel instantiation range listed in csrc/inst_groups.h becomes one object file
(inst.cu compiled with different -D flags) so the unrolled FFT kernels compile in parallel.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpragma_b200.so")
NVCC = os.environ.get("PDSP_NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
EXTRA = os.environ.get("PDSP_EXTRA_NVCC_FLAGS", "").split()  # experiments only (e.g. -DPDSP_F32_P32_FROM=7)
# --tuning / PDSP_TUNING=1: also build the 17 alternative N=1024 mappings (fft_config.h PDSP_VARIANT, selected with
# pdsp_ctx_tune("variant") / PDSP_VARIANT).  The shipped library carries the default mapping only.
TUNING = os.environ.get("PDSP_TUNING", "0") == "1"
LIB = os.environ.get("PDSP_LIB_OUT", LIB)  # experiment builds go to another file (and object directory)
OBJ = os.environ.get("PDSP_OBJ_DIR", OBJ)
FLAGS = ARCH + EXTRA + ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "--compress-mode=size", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                "-ccbin", "/usr/bin/g++"]


def groups():
    txt = open(os.path.join(CSRC, "inst_groups.h")).read()
    txt = txt[txt.index("#else"):]  # the list before #else is the reduced one of the emulated test build
    out = []
    for m in re.finditer(r"X\((\d), (\w+), (\w+), (\d+), (\d+)\)", txt):
        kind, tag, ctype, lo, hi = m.groups()
        if tag == "typetag":
            continue
        name = f"launch_{'r2c' if kind == '0' else 'c2c'}_{tag}_{lo}_{hi}"
        out.append((int(kind), ctype, int(lo), int(hi), name))
    return out


def _digest(paths, extra):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(open(p, "rb").read())
    return h.hexdigest()


def _compile(args):
    src, obj, flags, stamp = args
    cmd = [NVCC] + flags + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return obj, r.returncode, r.stdout + r.stderr
    open(obj + ".stamp", "w").write(stamp)
    return obj, 0, r.stderr


def build(force: bool = False, jobs: int | None = None, verbose: bool = True, tuning: bool | None = None) -> str:
    tuning = TUNING if tuning is None else tuning
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "pragma_b200.h"))
    tasks = []
    objs = []
    units = [(os.path.join(CSRC, "pragma_b200.cu"), os.path.join(OBJ, "pragma_b200.o"), []),
             (os.path.join(CSRC, "bigfft.cu"), os.path.join(OBJ, "bigfft.o"), []),
             (os.path.join(CSRC, "bigfft2.cu"), os.path.join(OBJ, "bigfft2.o"), []),
             (os.path.join(CSRC, "bigfft3.cu"), os.path.join(OBJ, "bigfft3.o"), [])]
    for kind, ctype, lo, hi, name in groups():
        defs = [f"-DPDSP_INST_KIND={kind}", f"-DPDSP_INST_T={ctype}", f"-DPDSP_INST_LO={lo}", f"-DPDSP_INST_HI={hi}",
                f"-DPDSP_INST_NAME={name}"]
        units.append((os.path.join(CSRC, "inst.cu"), os.path.join(OBJ, name + ".o"), defs))
    nvar = int(re.search(r"kNumVariants = (\d+)", open(os.path.join(CSRC, "fft_config.h")).read()).group(1))
    flags = FLAGS + (["-DPDSP_TUNING=1"] if tuning else [])
    for v in range(1, nvar if tuning else 1):
        for tag, ctype in (("f64", "double"), ("f32", "float")):
            name = f"launch_r2c_var_{tag}_{v}"
            units.append((os.path.join(CSRC, "inst_var.cu"), os.path.join(OBJ, name + ".o"),
                          [f"-DPDSP_VAR_T={ctype}", f"-DPDSP_VAR={v}", f"-DPDSP_INST_NAME={name}"]))
    for src, obj, defs in units:
        stamp = _digest(headers + [src], " ".join(flags + defs))
        objs.append(obj)
        old = open(obj + ".stamp").read() if os.path.exists(obj + ".stamp") and os.path.exists(obj) else ""
        if force or old != stamp:
            tasks.append((src, obj, flags + defs, stamp))
    if tasks:
        if verbose:
            print(f"[pragma_dsp_b200.build] compiling {len(tasks)} unit(s) for sm_100a ...", flush=True)
        with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            for obj, rc, log in ex.map(_compile, tasks):
                if rc != 0:
                    sys.stderr.write(log)
                    raise RuntimeError(f"nvcc failed for {obj}")
                if verbose and log.strip():
                    print(log.strip())
    stale = [o for o in os.listdir(OBJ) if o.endswith(".o") and os.path.join(OBJ, o) not in objs]
    for o in stale:  # objects of units no longer built (e.g. tuning variants after a default build)
        os.remove(os.path.join(OBJ, o))
    if tasks or stale or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-cudart", "static", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        if verbose:
            print(f"[pragma_dsp_b200.build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--tuning", action="store_true", help="also build the N=1024 tuning variants (-DPDSP_TUNING)")
    a = ap.parse_args()
    build(force=a.force, jobs=a.jobs, tuning=a.tuning or None)
