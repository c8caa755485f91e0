#!/usr/bin/env python3
"""Summarise an .ncu-rep (first kernel) into the handful of numbers DESIGN.md / profiles/ cite.

    python scripts/ncu_summary.py gpurun_out/prof_c2.ncu-rep [frames_per_launch]
"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def f(name):
    v = m.get(name, ("nan", ""))[0].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return float("nan")


def scaled(name):  # bytes with unit prefix -> bytes
    v, u = m.get(name, ("nan", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return float(v.replace(",", "")) * mult


cyc = f("sm__cycles_elapsed.avg")
sms = 148
s = {
    "kernel": m["Kernel Name"][0] if "Kernel Name" in m else None,
    "duration_us": f("gpu__time_duration.sum"),
    "dram_read_bytes": scaled("dram__bytes_read.sum"),
    "dram_write_bytes": scaled("dram__bytes_write.sum"),
    "dram_pct_of_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "registers_per_thread": f("launch__registers_per_thread"),
    "grid": f("launch__grid_size"), "block": f("launch__block_size"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "pipe_fp64_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "pipe_fma_pct": f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "pipe_alu_pct": f("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
    "lsu_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
    "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_active"),
    "shared_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "shared_bank_conflicts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "warp_insts": f("smsp__inst_executed.sum"),
    "sm_cycles": cyc,
    "per_frame": {
        "warp_insts": f("smsp__inst_executed.sum") / frames,
        "sm_cycles": cyc * sms / frames,
        "lsu_wavefronts": f("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed") / 100 * cyc * sms / frames,
        "shared_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / frames,
        "dram_bytes": (scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")) / frames,
    },
    "stalls_per_issue": {k.split("stalled_")[1].split("_per_issue")[0]: f(k) for k in hdr
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and f(k) > 0.05},
}
print(json.dumps(s, indent=1))
