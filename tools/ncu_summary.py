#!/usr/bin/env python
"""Summarise ncu artefacts into markdown for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_c3.csv
    python tools/ncu_summary.py kernel   gpurun_out/prof_tile_r1.ncu-rep
"""
import csv, io, subprocess, sys
from collections import defaultdict


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[hdr + 2:]:
        if len(r) > vi:
            d[r[ki].split("(")[0].replace("void ", "")].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) / len(v) for v in d.values())
    print("| kernel | launches | mean us | min us | max us | share of frame |\n|---|---|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
        m = sum(v) / len(v)
        print(f"| {k} | {len(v)} | {m / 1e3:.2f} | {min(v) / 1e3:.2f} | {max(v) / 1e3:.2f} | {100 * m / tot:.1f}% |")
    print(f"\nsum of per-kernel means: {tot / 1e3:.1f} us (ncu serialises launches and runs them cold; compare shares, not absolutes)")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for vals in rows[2:]:
        if len(vals) == len(h):
            kernel_row(h, units, vals)
            print()


def kernel_row(h, units, vals):
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_static",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
    print(f"kernel: {vals[h.index('Kernel Name')]}\n\n| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in h:
            i = h.index(w)
            print(f"| {w} | {vals[i]} | {units[i]} |")
    stalls = [(float(vals[i] or 0), n) for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")]
    if stalls:
        print("\ntop warp stall reasons (warps stalled per issue-active cycle):\n")
        for v, n in sorted(stalls, reverse=True)[:6]:
            print(f"- {n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
