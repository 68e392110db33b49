# round 2: N-GPU lines (usage: gpu_r2_multi.sh N TAG): the default bench (standing_4096 per GPU + also-lines trot_65536 and the 131072-per-GPU
# shard of the 1 M-instance config) and BASELINE config 5 (262144-instance push sweep, 400 closed-loop cycles, strong scaling)
N=$1; TAG=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --workload push_sweep --sweep-cycles 400 > gpurun_out/${TAG}_sweep_${N}gpu.json 2> gpurun_out/${TAG}_sweep_${N}gpu.err
python - <<PY
import json
for f in ("${TAG}_bench_${N}gpu","${TAG}_sweep_${N}gpu"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f n_gpus %s ms/step %.3f" % (d["value"], d["e2e"]["value"], d["n_gpus"], d["ms_per_step"]))
        for k,v in (d.get("also") or {}).items(): print("   also", k, "value %.0f ms/step %.3f" % (v["value"], v["ms_per_step"]))
        if "sweep" in d: print("   ", json.dumps(d["sweep"])[:1500])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
