#!/usr/bin/env python3
"""Digest of an ncu source page (ncu -i x.ncu-rep --page source --csv): stall-reason totals, the instructions with the most
stall samples, and shared-memory wavefront excess per instruction.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python scripts/ncu_source_digest.py src.csv [top]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
# the file holds one table per captured kernel, each preceded by a 'Kernel Name' row
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        tables.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for t in tables:
    h = {n: i for i, n in enumerate(t["hdr"])}
    print("==", t["name"][:110])
    stalls = [n for n in t["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    samples = 0
    for r in t["rows"]:
        samples += int(r[h["# Samples"]] or 0)
        for s in stalls:
            tot[s] += int(r[h[s]] or 0)
    print("  samples", samples, " ".join(f"{k[6:]}={100 * v / max(1, sum(tot.values())):.1f}%" for k, v in tot.most_common(9)))
    ex = sum(int(r[h["L1 Wavefronts Shared Excessive"]] or 0) for r in t["rows"])
    sh = sum(int(r[h["L1 Wavefronts Shared"]] or 0) for r in t["rows"])
    print(f"  shared wavefronts {sh}, excessive {ex} ({100 * ex / max(1, sh):.1f}%)")
    by = sorted(t["rows"], key=lambda r: -int(r[h["# Samples"]] or 0))[:top]
    for r in by:
        why = max(stalls, key=lambda s: int(r[h[s]] or 0))
        print(f"   {int(r[h['# Samples']]):6d}  {r[h['Source']].strip()[:70]:70s} {why[6:]}  exc_wf={r[h['L1 Wavefronts Shared Excessive']]}")
    worst = sorted(t["rows"], key=lambda r: -int(r[h["L1 Wavefronts Shared Excessive"]] or 0))[:6]
    for r in worst:
        if int(r[h["L1 Wavefronts Shared Excessive"]] or 0):
            print(f"   conflict: {r[h['Source']].strip()[:60]:60s} wf={r[h['L1 Wavefronts Shared']]} ideal={r[h['L1 Wavefronts Shared Ideal']]}")
