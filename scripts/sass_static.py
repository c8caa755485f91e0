#!/usr/bin/env python3
"""Static opcode histogram of one kernel in an object file (cuobjdump -sass).  The frame loop of the r2c kernels is
fully unrolled straight-line code, so static counts ~ per-frame counts (plus prologue / finishing loop).

    python scripts/sass_static.py file.o 'r2c_kernelIdLi9ELi4ELi4ELi128ELi4ELi1E' [top]
"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, mix = None, collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or pat not in cur:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(2)
        base = op.split(".")[0]
        key = op if base in ("LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "LDL", "STL", "UBLKCP", "SYNCS") else base
        mix[key] += 1
tot = sum(mix.values())
print(f"{pat}: {tot} static instructions")
dp = sum(v for k, v in mix.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
print(f"  DP pipe {dp}")
for k, v in mix.most_common(top):
    print(f"  {k:24s} {v}")
