timeout 900 python bench.py --workload push_sweep --per-gpu 32768 --sweep-cycles 400 --plant dynamics > gpurun_out/r2aj_sweep_dyn.json 2> gpurun_out/r2aj.err; tail -2 gpurun_out/r2aj.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2aj_sweep_dyn.json").read().strip().splitlines()[-1])
print("value %.0f ms/step %.3f failures %s" % (d["value"], d["ms_per_step"], d["stats"]["solver_failures"]))
for g,v in d["stats"]["sweep"]["per_gain"].items(): print(g, {k:(round(x,4) if isinstance(x,float) else x) for k,x in v.items() if k!="first_order_prediction"})
print({k:v for k,v in d["stats"].items() if k!="sweep"})
PY
