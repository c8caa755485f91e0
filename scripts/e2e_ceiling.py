#!/usr/bin/env python3
"""Host <-> device copy ceiling of the box: G concurrent streams of pinned H2D + D2H copies, no kernel.

    python scripts/e2e_ceiling.py [--mb 512] [--seconds 0.6]

For G = 1, 2, 4, 8 (up to the device count): one host thread per device issues the north-star job's traffic pattern (2 bytes
in per byte out, H2D and D2H on separate streams so both DMA directions run) from pinned buffers for `seconds`; prints the
aggregate GB/s per direction and the frames/s of the north-star shape (8192 B in + 4104 B out per frame) that rate allows.
Torch is measurement tooling here (pinned tensors, streams); the product is not involved."""
import argparse
import json
import threading
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=512)
ap.add_argument("--seconds", type=float, default=0.6)
ap.add_argument("--numa", action="store_true", help="allocate each device's pinned buffers from a thread pinned to the device's NUMA node")
a = ap.parse_args()


def numa_cpus(d):
    import os
    p = torch.cuda.get_device_properties(d)
    bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for tok in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = tok.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        return set(cpus) & os.sched_getaffinity(0) or None
    except Exception:
        return None
ndev = torch.cuda.device_count()
rows = []
for G in (1, 2, 4, 8):
    if G > ndev:
        break
    bufs = [None] * G

    def alloc(d):
        import os
        if a.numa:
            cpus = numa_cpus(d)
            if cpus:
                os.sched_setaffinity(0, cpus)
        hin = torch.empty(a.mb << 20, dtype=torch.uint8).pin_memory()
        hout = torch.empty((a.mb << 20) // 2, dtype=torch.uint8).pin_memory()
        hin.fill_(1)
        hout.fill_(1)
        din = torch.empty(a.mb << 20, dtype=torch.uint8, device=f"cuda:{d}")
        dout = torch.empty((a.mb << 20) // 2, dtype=torch.uint8, device=f"cuda:{d}")
        bufs[d] = (hin, hout, din, dout, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d))

    ath = [threading.Thread(target=alloc, args=(d,)) for d in range(G)]
    for t in ath:
        t.start()
    for t in ath:
        t.join()
    counts = [0] * G
    start = threading.Barrier(G + 1)
    stop = threading.Event()

    def worker(d):
        hin, hout, din, dout, s1, s2 = bufs[d]
        torch.cuda.set_device(d)
        start.wait()
        while not stop.is_set():
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
            counts[d] += 1

    th = [threading.Thread(target=worker, args=(d,)) for d in range(G)]
    for t in th:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    time.sleep(a.seconds)
    stop.set()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    h2d = sum(counts) * (a.mb << 20) / dt / 1e9
    d2h = h2d / 2
    row = {"gpus": G, "numa_local_buffers": bool(a.numa), "h2d_gbs": h2d, "d2h_gbs": d2h, "north_star_frames_per_s_ceiling": h2d * 1e9 / 8192,
           "per_gpu_h2d_gbs": h2d / G}
    rows.append(row)
    print(json.dumps(row), flush=True)
    del bufs
