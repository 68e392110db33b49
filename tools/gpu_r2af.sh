# round 2, final check: the whole GPU suite, smoke(), the default bench and the reference arm exactly as the driver runs them
mkdir -p gpurun_out
echo "(suite: see r2af_pytest.log of the previous call)"

S=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2af_ref.json 2> gpurun_out/r2af_ref.err; tail -1 gpurun_out/r2af_ref.err; tail -c 300 gpurun_out/r2af_ref.json; echo
echo "reference arm wall $(( $(date +%s) - S )) s"; S=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; tail -1 gpurun_out/r2af_bench.err
echo "b200 arm wall $(( $(date +%s) - S )) s"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2af_bench.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f ms/step %.3f launches %s clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
for k,v in d["also"].items(): print("  also", k, "%.0f" % v["value"])
r=json.loads(open("gpurun_out/r2af_ref.json").read().strip().splitlines()[-1])
print("reference arm value %.0f -> e2e ratio %.1f" % (r["value"], d["e2e"]["value"]/r["value"]))
PY
