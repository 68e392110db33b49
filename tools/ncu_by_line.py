#!/usr/bin/env python3
"""Executed warp instructions and stall samples per CUDA source line of one kernel (all inlined code attributed to the line it came
from; non-inlined device functions included), from an ncu report plus nvdisasm's line info of the library that was profiled.
    python tools/ncu_by_line.py REPORT.ncu-rep LIB.so KERNEL [top]"""
import csv, re, subprocess, sys, collections, tempfile, os
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# section of the kernel: map instruction offset -> (file, line, inlined-at chain's outermost line)
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern + "N" in l or l.startswith(".text._Z%d%s" % (len(kern), kern)))
off2line = {}
cur = ("?", 0)
fl = re.compile(r'//## File "([^"]+)", line (\d+)')
ins = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+\S")
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.lstrip().startswith(".section"): 
        if off2line: break
        continue
    m = fl.search(l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = ins.match(l)
    if m: off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:^" + kern + "$", "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
s0 = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
s1 = next((i for i in range(s0 + 1, len(out)) if out[i].startswith('"Kernel Name"')), len(out))
rows = list(csv.DictReader(out[s0:s1]))
base = int(rows[0]["Address"], 16)
ex = collections.Counter(); sm = collections.Counter()
for r in rows:
    o = int(r["Address"], 16) - base
    k = off2line.get(o, ("?", 0))
    ex[k] += int(r["Instructions Executed"] or 0); sm[k] += int(r["# Samples"] or 0)
te, ts = sum(ex.values()), sum(sm.values())
print("kernel %s: %d warp instructions, %d samples; top %d source lines by executed instructions" % (kern, te, ts, top))
for k, v in ex.most_common(top):
    print("%6.2f%% inst %6.2f%% samples  %s:%d" % (100.0 * v / te, 100.0 * sm[k] / max(ts, 1), k[0], k[1]))
byfile = collections.Counter()
for k, v in ex.items(): byfile[k[0]] += v
print("by file:", {f: "%.1f%%" % (100.0 * v / te) for f, v in byfile.most_common()})
