#!/usr/bin/env python3
"""Shared-memory bank-conflict model for FftEngine's exchange patterns (design aid).

For a kernel configuration (element bytes, LOG2M, LOG2P, radix bits per pass, pad unit) it
enumerates every warp-wide STS/LDS the engine issues and counts wavefronts under the usual
model: a 16-byte access is served in 4 phases of 8 lanes, an 8-byte access in 2 phases of 16
lanes; inside a phase the cost is the maximum number of distinct 4-byte-bank rows any bank sees.

    python scripts/bank_conflicts.py
"""
import itertools


def wavefronts(addrs_bytes, ebytes):
    lanes_per_phase = {16: 8, 8: 16, 4: 32}[ebytes]
    total = 0
    for p0 in range(0, len(addrs_bytes), lanes_per_phase):
        grp = addrs_bytes[p0:p0 + lanes_per_phase]
        banks = {}
        for a in grp:
            for w in range(ebytes // 4):
                b = ((a // 4) + w) % 32
                banks.setdefault(b, set()).add((a + 4 * w) // 128)
        total += max(len(v) for v in banks.values()) if banks else 0
    return total


def analyse(ebytes, log2m, log2p, rb, pad_unit, frame_stride_pad=0, verbose=False):
    M, P = 1 << log2m, 1 << log2p
    TF = M // P
    npass = -(-log2m // rb)
    bits = [rb] * (npass - 1) + [log2m - rb * (npass - 1)]
    pad = (lambda i: i + i // pad_unit) if pad_unit else (lambda i: i)
    slot_elems = M + (M // pad_unit if pad_unit else 0) + 1 + frame_stride_pad
    ideal = {16: 4, 8: 2, 4: 1}[ebytes]
    res = []
    nsl = 0
    for pi, b in enumerate(bits[:-1]):
        R = 1 << b
        bpt = P // R
        NS = 1 << nsl
        w_tot = w_n = 0
        for u, k in itertools.product(range(bpt), range(R)):
            for warp0 in range(0, max(TF, 32), 32):  # one warp = 32 consecutive tids (maybe several frames)
                addrs = []
                for lane in range(32):
                    tid = warp0 + lane
                    slot, t = divmod(tid, TF)
                    if TF >= 32 and slot > 0:
                        continue
                    j = t + TF * u
                    base = ((j >> nsl) << (nsl + b)) + (j & (NS - 1))
                    addrs.append((slot * slot_elems + pad(base + (k << nsl))) * ebytes)
                w_tot += wavefronts(addrs, ebytes)
                w_n += 1
                if TF < 32:
                    break
        r_tot = r_n = 0
        for q in range(P):
            for warp0 in range(0, max(TF, 32), 32):
                addrs = []
                for lane in range(32):
                    tid = warp0 + lane
                    slot, t = divmod(tid, TF)
                    if TF >= 32 and slot > 0:
                        continue
                    addrs.append((slot * slot_elems + pad(t + TF * q)) * ebytes)
                r_tot += wavefronts(addrs, ebytes)
                r_n += 1
                if TF < 32:
                    break
        res.append((pi, R, w_tot / w_n / ideal, r_tot / r_n / ideal))
        nsl += b
    return res


if __name__ == "__main__":
    print("elem  M     P   rb pad | per exchange: (pass, radix, write x-ideal, read x-ideal)")
    for ebytes in (16, 8):
        for log2m in (6, 7, 8, 9, 10, 11, 12):
            for log2p, rb in ((3, 3), (4, 3), (4, 4), (5, 5)):
                if log2p > log2m or -(-log2m // rb) < 2:
                    continue
                row = 128 // ebytes
                for pad_unit in sorted({row, max(row, 1 << rb), 0}):
                    r = analyse(ebytes, log2m, log2p, rb, pad_unit)
                    print(f"{ebytes:4d} {1 << log2m:5d} {1 << log2p:3d} {rb:3d} {pad_unit:3d} | " +
                          "  ".join(f"p{p} R{R} w{w:.2f} r{rd:.2f}" for p, R, w, rd in r))
