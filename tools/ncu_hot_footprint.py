#!/usr/bin/env python3
"""Hot instruction footprint of a kernel from an ncu --set full --import-source report: SASS instructions sorted by executed
count, cumulative bytes (16 B each) against the cumulative share of executed instructions; and the same per device function.
    python tools/ncu_hot_footprint.py REPORT.ncu-rep KERNEL_REGEX [LIB.so]"""
import csv, re, subprocess, sys
rep, kern = sys.argv[1:3]
lib = sys.argv[3] if len(sys.argv) > 3 else "wbc_quadruped_dob_b200/lib/libwbc_b200.so"
out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kern, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(lines[start:end]))
ex = sorted((int(r["Instructions Executed"] or 0) for r in rows), reverse=True)
tot = sum(ex)
print("kernel %s: %d SASS instructions (%.1f KB), %d executed (warp level)" % (kern, len(rows), len(rows) * 16 / 1024.0, tot))
acc = 0
marks = [0.5, 0.8, 0.9, 0.95, 0.99, 0.999]
mi = 0
for k, e in enumerate(ex):
    acc += e
    while mi < len(marks) and acc >= marks[mi] * tot:
        print("  %5.1f%% of executed instructions come from the hottest %6.1f KB of code" % (100 * marks[mi], (k + 1) * 16 / 1024.0))
        mi += 1
nz = sum(1 for e in ex if e > 0)
print("  instructions executed at least once: %.1f KB" % (nz * 16 / 1024.0))
# per function (needs the ELF symbol table of the same build)
try:
    elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
    funcs = []
    for line in elf.splitlines():
        m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", line)
        if m and re.search(kern, m.group(3)):
            funcs.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
    base = int(rows[0]["Address"], 16)
    print("  per function: static KB | KB executed at all | KB holding 95%% of the function's executed instructions | share of kernel's executed instructions")
    res = []
    for off, size, name in funcs:
        sel = sorted((int(r["Instructions Executed"] or 0) for r in rows if off <= int(r["Address"], 16) - base < off + size), reverse=True)
        t = sum(sel)
        if t == 0: continue
        a, k95 = 0, 0
        for k, e in enumerate(sel):
            a += e
            if a >= 0.95 * t: k95 = k + 1; break
        res.append((t, name, size, sum(1 for e in sel if e > 0), k95))
    for t, name, size, nzf, k95 in sorted(res, reverse=True)[:28]:
        print("   %6.1f | %6.1f | %6.1f | %5.1f%%  %s" % (size / 1024.0, nzf * 16 / 1024.0, k95 * 16 / 1024.0, 100.0 * t / tot, name[:70]))
except Exception as e:
    print("  (per-function table unavailable: %s)" % e)
