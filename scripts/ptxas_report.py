#!/usr/bin/env python3
"""Compile one instantiation unit with -Xptxas -v and print registers / spills per kernel.

    python scripts/ptxas_report.py r2c double 9 9 [extra nvcc flags]     # kind type lo hi
"""
import re
import subprocess
import sys

kind, ctype, lo, hi = sys.argv[1:5]
extra = sys.argv[5:]
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
       "--expt-relaxed-constexpr", "-Xptxas", "-v", f"-DPDSP_INST_KIND={0 if kind == 'r2c' else 1}", f"-DPDSP_INST_T={ctype}",
       f"-DPDSP_INST_LO={lo}", f"-DPDSP_INST_HI={hi}", "-DPDSP_INST_NAME=unit", "-c", "pragma_dsp_b200/csrc/inst.cu",
       "-o", "/tmp/ptxas_report.o"] + extra
out = subprocess.run(cmd, capture_output=True, text=True)
txt = out.stdout + out.stderr
if out.returncode:
    print(txt)
    sys.exit(1)
cur = None
rows = {}
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        rows[cur] = {}
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and "spill" not in rows[cur]:
        rows[cur]["stack"], rows[cur]["spill"] = int(m.group(1)), (int(m.group(2)), int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        rows[cur]["regs"] = int(m.group(1))
for k, v in rows.items():
    k = re.sub(r"void pdsp::|\(pdsp::\w+\)", "", k)
    print(f"{k:70s} regs {v.get('regs'):4d}  stack {v.get('stack', 0):4d}  spill st/ld {v.get('spill')}")
