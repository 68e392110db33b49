#!/usr/bin/env python3
"""Turn the outputs of tools/gpu_r2x.sh (gpurun_out/r2x_*) into the committed evidence under profiles/ (r02_x_*), refresh
profiles/ncu_traffic.json from the two --set full captures, and print the numbers DESIGN.md quotes."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
def run(cmd, out):
    with open(out, "w") as f:
        subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, cwd=ROOT)
for tag, kern in (("4096", "wbc_solve_kernel"), ("65536", "wbc_solve_staged_kernel")):
    rep = os.path.join(G, "r2x_full_%s.ncu-rep" % tag)
    run([sys.executable, "tools/ncu_summary.py", rep], os.path.join(P, "r02_x_%s_ncu_summary.txt" % tag))
    run([sys.executable, "tools/ncu_by_function.py", rep, os.path.join(G, "r2x_lib.so"), kern], os.path.join(P, "r02_x_%s_by_function.txt" % tag))
    run([sys.executable, "tools/ncu_hot_footprint.py", rep, kern], os.path.join(P, "r02_x_%s_hot_footprint.txt" % tag))
for f in ("bench_default", "bench_ref", "bench_rollout_4096", "bench_rollout_65536", "bench_sweep_32768", "bench_single"):
    shutil.copy(os.path.join(G, "r2x_%s.json" % f), os.path.join(P, "r02_x_%s.json" % f))
shutil.copy(os.path.join(G, "r2x_launches.csv"), os.path.join(P, "r02_x_launches.csv"))
# traffic per solver launch from the raw pages
traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE solver launch from the ncu --set full capture named in `source` (bench.py copies this into roofline.traffic for the matching workload)"}
for tag, kern, wl in (("4096", "wbc_solve_kernel", "standing_4096"), ("65536", "wbc_solve_staged_kernel", "trot_65536")):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, "r2x_full_%s.ncu-rep" % tag), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]; units = rows[1]
    for r in rows[2:]:
        if len(r) < len(h) or not r[h.index("Kernel Name")].startswith(kern + "("): continue
        def val(name):
            i = h.index(name); v = float(r[i]); u = units[i]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        traffic[wl] = {"bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr), "kernel": kern, "source": "profiles/r02_x_%s_ncu_summary.txt" % tag}
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
d = json.loads(open(os.path.join(G, "r2x_bench_default.json")).read().strip().splitlines()[-1])
def line(x):
    r = x["roofline"]
    return "value %.3f M  e2e %.3f M  front %.3f + solve %.2f ms  fifo %.3f M  frac %.4f" % (x["value"] / 1e6, x["e2e"]["value"] / 1e6, r["front_kernel_ms"], r["kernel_ms"], (x.get("value_fifo") or 0) / 1e6, r["frac"])
print("standing_4096 ", line(d))
for k, v in d["also"].items(): print(k, line(v))
print("cpu_baseline", d["cpu_baseline"])
print(json.dumps(traffic, indent=1))
