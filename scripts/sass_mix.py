#!/usr/bin/env python3
"""Opcode histogram (warp-level executed instructions) from an ncu source page.

    ncu -i x.ncu-rep --page source --csv > src.csv ; python scripts/sass_mix.py src.csv [frames]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
hdr = rows[1]
i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
i_samp = hdr.index("# Samples")
mix, samp = collections.Counter(), collections.Counter()
total = 0
for r in rows[2:]:
    if len(r) <= i_exec:
        continue
    toks = r[i_src].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.rstrip(";")
    base = op.split(".")[0]
    key = op if base in ("LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "LDC", "LDL", "STL") else base
    n = int(r[i_exec] or 0)
    mix[key] += n
    samp[key] += int(r[i_samp] or 0)
    total += n
print(f"total warp instructions {total}  ({total / frames:.1f} per frame)")
tot_s = sum(samp.values()) or 1
for k, n in mix.most_common(40):
    print(f"{k:28s} {n / frames:9.1f} /frame  {100 * n / total:5.1f}%   stall samples {100 * samp[k] / tot_s:5.1f}%")
