#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by device function (the noinline callees of a kernel).

    python tools/ncu_by_function.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING

Prints, per function: executed warp instructions, share, stall samples, top stall reasons."""
import csv
import re
import subprocess
import sys


def main():
    rep, lib, kern = sys.argv[1:4]
    elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
    funcs = []
    for line in elf.splitlines():
        m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", line)
        if m and kern in m.group(3):
            funcs.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
    funcs.sort()
    out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kern, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))   # first launch only
    rd = csv.DictReader(lines[start:end])
    rows = list(rd)
    base = int(rows[0]["Address"], 16)
    stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    agg = {}
    for r in rows:
        off = int(r["Address"], 16) - base
        name = "<kernel body>"
        for a, s, n in funcs:
            if a <= off < a + s:
                name = n
                break
        d = agg.setdefault(name, {"inst": 0, "samples": 0, "stalls": {}})
        d["inst"] += int(r["Instructions Executed"] or 0)
        d["samples"] += int(r["# Samples"] or 0)
        for c in stall_cols:
            d["stalls"][c] = d["stalls"].get(c, 0) + int(r[c] or 0)
    tot_i = sum(d["inst"] for d in agg.values())
    tot_s = sum(d["samples"] for d in agg.values())
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        top = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:4]
        short = re.sub(r"INS_6TeamExILi\d+EEE.*", "", name)
        noi = sum(v for k, v in d["stalls"].items() if "no_inst" in k)
        print("%-60s inst %5.1f%%  samples %5.1f%%  no_inst(all) %4.2f%%  %s" % (short[:60], 100.0 * d["inst"] / max(tot_i, 1), 100.0 * d["samples"] / max(tot_s, 1), 100.0 * noi / max(tot_s, 1),
                                                         " ".join("%s=%.0f%%" % (k[6:], 100.0 * v / max(d["samples"], 1)) for k, v in top)))


if __name__ == "__main__":
    main()
