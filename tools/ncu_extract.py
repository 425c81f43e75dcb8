#!/usr/bin/env python
"""Turn `ncu --set full` captures into the small JSON bench.py reads for `roofline.traffic` and the L2-side figure.

    python tools/ncu_extract.py profiles/r2_ncu_kernels.json "<command that was profiled>" gpurun_out/prof_bwd.ncu-rep gpurun_out/prof_fwd.ncu-rep

Per kernel (averaged over the captured launches): dram__bytes_read.sum + dram__bytes_write.sum, lts__t_sectors.sum * 32 B,
gpu__time_duration.sum and a few occupancy / issue counters.  The file records the git revision it was produced on, so a
bench line can say which build its `traffic` figure belongs to.  Run in the build container (ncu reads reports without a GPU).
"""
import csv
import json
import subprocess
import sys

WANT = {
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "lts__t_sectors.sum": "lts_sectors",
    "gpu__time_duration.sum": "duration", "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__cycles_active.avg": "cycles_active", "sm__cycles_elapsed.max": "cycles_elapsed", "launch__waves_per_multiprocessor": "waves",
    "launch__registers_per_thread": "registers", "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum": "red_requests", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "ld_requests",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed": "l2_atomic_pct",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}


def num(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def main():
    out_path, command, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    kernels = {}
    for rep in reps:
        text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(text.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, data = rows[0], rows[1], rows[2:]
        name_i = hdr.index("Kernel Name")
        for r in data:
            rec = kernels.setdefault(r[name_i], {"launches": 0, "report": rep.split("/")[-1]})
            rec["launches"] += 1
            for i, h in enumerate(hdr):
                if h in WANT and num(r[i]) is not None:
                    rec.setdefault("_" + WANT[h], []).append(num(r[i]) * SCALE.get(units[i], 1.0))
    for rec in kernels.values():
        for key in [k for k in rec if k.startswith("_")]:
            vals = rec.pop(key)
            rec[key[1:]] = sum(vals) / len(vals)
        if "dram_read" in rec:
            rec["dram_bytes"] = rec["dram_read"] + rec.get("dram_write", 0.0)
        if "lts_sectors" in rec:
            rec["lts_bytes"] = rec["lts_sectors"] * 32.0
        if "duration" in rec:
            rec["duration_us"] = rec.pop("duration")
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"git": git, "command": command, "note": "per-launch averages; times under ncu are cold-cache and serialised -- never bench values",
               "kernels": kernels}, open(out_path, "w"), indent=1)
    for name, rec in kernels.items():
        print(name[:80], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in rec.items()})


if __name__ == "__main__":
    main()
