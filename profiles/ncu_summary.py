#!/usr/bin/env python
"""Summarise ncu artefacts brought back from gpurun (run in the build container, no GPU needed).

    python profiles/ncu_summary.py launches gpurun_out/launches.csv        # per-kernel share of a step
    python profiles/ncu_summary.py raw gpurun_out/prof.ncu-rep [substr..]  # selected raw metrics of a full capture
"""
import collections
import csv
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_global.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1 :]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg[r[ki][:90]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':90s} {'n':>5s} {'mean us':>9s} {'total us':>10s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:90s} {len(v):5d} {sum(v)/len(v):9.2f} {sum(v):10.1f} {100*sum(v)/tot:5.1f}%")


def raw(path, substrs):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = hdr.index("Kernel Name")
    for r in data:
        print("kernel:", r[names][:100])
    for i, h in enumerate(hdr):
        if h in KEY_METRICS or any(s in h for s in substrs):
            print(f"{h:95s} {units[i]:14s} {[r[i] for r in data]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3:])
