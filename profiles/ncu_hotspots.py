#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples of one kernel in an ncu report (needs -lineinfo / --import-source on).

    python profiles/ncu_hotspots.py gpurun_out/prof.ncu-rep [N]
"""
import csv
import subprocess
import sys


def main(path, n=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    print(rows[hi - 1][:2])
    hdr = rows[hi]
    si, wi, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    first, seen = [], set()
    for r in rows[hi + 1 :]:
        if len(r) != len(hdr) or r[0] in seen:
            break
        seen.add(r[0])
        first.append(r)
    tot = sum(int(r[wi] or 0) for r in first)
    toti = sum(int(r[ii] or 0) for r in first)
    print(f"{len(first)} SASS instructions, {tot} stall samples, {toti} warp-instructions executed")
    ops = {}
    for r in first:
        op = r[si].split()[0] if not r[si].startswith("@") else r[si].split()[1]
        op = op.split(".")[0]
        e = ops.setdefault(op, [0, 0])
        e[0] += int(r[ii] or 0)
        e[1] += int(r[wi] or 0)
    print("by opcode (executed share / stall share):")
    for op, (e, w) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]:
        print(f"  {op:10s} {100*e/toti:5.1f}%  {100*w/tot:5.1f}%")
    print("top instructions by stall samples:")
    for idx, r in sorted(enumerate(first), key=lambda t: -int(t[1][wi] or 0))[:n]:
        print(f"  #{idx:5d} {int(r[wi]):5d} {100*int(r[wi])/tot:5.1f}%  exec={int(r[ii] or 0):8d}  {r[si][:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
