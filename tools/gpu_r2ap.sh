# final library, N GPUs (usage: gpu_r2ap.sh N): the default bench as the driver launches it (headline standing_4096 per GPU + also-lines), no CPU baseline
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG:-r2ap}_bench_${N}gpu.json 2> gpurun_out/${TAG:-r2ap}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG:-r2ap}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f n_gpus %s ms/step %.3f p50 %.3f" % (d["value"], d["e2e"]["value"], d["n_gpus"], d["ms_per_step"], d["p50_ms"]))
for k,v in (d.get("also") or {}).items(): print("   also", k, "value %.0f e2e %.0f ms/step %.3f" % (v["value"], v["e2e"]["value"], v["ms_per_step"]))
PY
tail -2 gpurun_out/${TAG:-r2ap}_bench_${N}gpu.err
