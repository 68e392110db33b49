#!/usr/bin/env python3
"""Code-size attribution: SASS instructions per source file:line (and per enclosing function symbol) of one kernel.
usage: sass_by_line.py all.sass <start-label-substring> [end-label-substring]   (all.sass = nvdisasm --print-line-info)"""
import re, sys, collections
path, start = sys.argv[1], sys.argv[2]
end = sys.argv[3] if len(sys.argv) > 3 else None
lines = open(path).read().splitlines()
i0 = next(i for i, l in enumerate(lines) if l.startswith(start))
i1 = len(lines)
for i in range(i0 + 1, len(lines)):
    l = lines[i]
    if (end and l.startswith(end)) or (not end and (l.startswith("$_Z") or l.startswith(".text.") or l.startswith("_Z")) and i > i0 + 2):
        i1 = i; break
cur = ("?", 0); byline = collections.Counter(); inl = collections.Counter()
inst = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S")
fl = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
n = 0
for l in lines[i0:i1]:
    m = fl.search(l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if inst.match(l):
        byline[cur] += 1; n += 1
print("instructions:", n, "bytes:", n * 16)
byfile = collections.Counter()
for (f, ln), c in byline.items(): byfile[f] += c
print("by file:", dict(byfile))
# bucket by 20-line windows for readability
buck = collections.Counter()
for (f, ln), c in byline.items(): buck[(f, ln // 10 * 10)] += c
for (f, ln), c in buck.most_common(45): print("%6d  %s:%d" % (c, f, ln))
