#!/usr/bin/env python3
"""SASS bytes per device function of one kernel in libwbc_b200.so (cuobjdump -elf symbol table)."""
import subprocess, sys, re
lib = sys.argv[1] if len(sys.argv) > 1 else "wbc_quadruped_dob_b200/lib/libwbc_b200.so"
kern = sys.argv[2] if len(sys.argv) > 2 else "wbc_solve_kernel"
out = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rows = []
for l in out.splitlines():
    p = l.split()
    if len(p) >= 7 and p[0].startswith("0x") and kern in p[-1] and "$" in p[-1]:
        try:
            rows.append((int(p[2], 16), p[-1].split("$")[-1][:80]))
        except ValueError:
            pass
    elif len(p) >= 7 and p[0].startswith("0x") and p[-1].startswith("_Z") and kern in p[-1]:
        total = int(p[2], 16)
rows.sort(reverse=True)
sub = sum(s for s, _ in rows)
print("kernel total %d B, kernel body %d B" % (total, total - sub))
for s, n in rows: print("%7d  %s" % (s, n))
