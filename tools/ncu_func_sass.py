#!/usr/bin/env python3
"""Per-instruction executed counts / stall samples of one device function from an ncu report.
    python tools/ncu_func_sass.py REPORT.ncu-rep LIB.so KERNEL FUNC_SUBSTRING [min_exec_share]"""
import csv, re, subprocess, sys
rep, lib, kern, fn = sys.argv[1:5]
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rng = None
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", line)
    if m and kern in m.group(3) and fn in m.group(4):
        rng = (int(m.group(1), 16), int(m.group(2), 16)); break
out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kern, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(lines[start:end]))
base = int(rows[0]["Address"], 16)
sel = [r for r in rows if rng[0] <= int(r["Address"], 16) - base < rng[0] + rng[1]]
tot = sum(int(r["Instructions Executed"] or 0) for r in sel); ts = sum(int(r["# Samples"] or 0) for r in sel)
print("function %s: %d SASS instructions, %d executed, %d samples" % (fn, len(sel), tot, ts))
for r in sel:
    e = int(r["Instructions Executed"] or 0); s = int(r["# Samples"] or 0)
    print("%5x %9d %5.1f%% s%5.1f%%  %s" % (int(r["Address"], 16) - base - rng[0], e, 100.0 * e / max(tot, 1), 100.0 * s / max(ts, 1), r["Source"][:70]))
