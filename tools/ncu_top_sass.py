#!/usr/bin/env python3
"""Top SASS instructions of a kernel by stall samples, with the owning device function.
    python tools/ncu_top_sass.py REPORT.ncu-rep LIB.so KERNEL [N]"""
import csv, re, subprocess, sys
rep, lib, kern = sys.argv[1:4]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
funcs = []
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", line)
    if m and kern in m.group(3):
        funcs.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kern, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(lines[start:end]))
base = int(rows[0]["Address"], 16)
tot = sum(int(r["# Samples"] or 0) for r in rows)
def fn(off):
    for a, s, n in funcs:
        if a <= off < a + s: return re.sub(r"^_ZN5wbcqp(4fast)?\d+", "", n)[:28]
    return "<body>"
rows.sort(key=lambda r: -int(r["# Samples"] or 0))
for r in rows[:N]:
    off = int(r["Address"], 16) - base
    print("%5.2f%%  %-28s %6x  %s" % (100.0 * int(r["# Samples"] or 0) / tot, fn(off), off, r["Source"][:60]))
