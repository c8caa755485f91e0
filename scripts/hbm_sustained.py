#!/usr/bin/env python3
"""CONTEXT: what a plain device-to-device copy sustains on this box - best single call (the way MEASURED_PEAKS.json's hbm_gbs
was taken: best of 10) against back-to-back calls for ~0.3 s and ~2 s (power cap), with the clocks sampled meanwhile.  Also a
2:1 read:write stream (the north-star kernel's mix: 8 KB in, 4 KB out per frame) built from torch ops."""
import json
import subprocess
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

dev = torch.device("cuda", 0)
n = 1 << 30  # bf16 elements: 2 GiB read + 2 GiB written per copy
a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)


def clocks():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return [float(x) for x in out.split(",")]
    except Exception:
        return None


def run(fn, nbytes, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        time.sleep(0.05)
    reps = max(10, int(seconds * 1e3 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    ck = clocks()
    torch.cuda.synchronize()
    return nbytes / (best * 1e-3) / 1e9, nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, ck


for secs in (0.3, 2.0):
    burst, sus, ck = run(lambda: b.copy_(a), 4.0 * n, secs)
    print(json.dumps({"what": "copy 2 GiB -> 2 GiB (1:1)", "best_single_gbs": burst, "sustained_gbs": sus, "seconds": secs, "clocks_sm_mem_power": ck}), flush=True)
# 2:1 read:write: out[i] = x[2i] + x[2i+1] (reads 2 bytes per byte written)
x = a.view(torch.float32)          # 2^29 floats (2 GiB)
y = torch.empty(x.numel() // 2, dtype=torch.float32, device=dev)
x2 = x.view(-1, 2)
for secs in (0.3, 2.0):
    burst, sus, ck = run(lambda: torch.sum(x2, dim=1, out=y), 4.0 * x.numel() * 1.5, secs)
    print(json.dumps({"what": "pair sum (2:1 read:write), 2 GiB in -> 1 GiB out", "best_single_gbs": burst, "sustained_gbs": sus, "seconds": secs, "clocks_sm_mem_power": ck}), flush=True)
