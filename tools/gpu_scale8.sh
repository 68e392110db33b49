# 8-GPU weak-scaling lines: the default workload, the 65 536-per-GPU trot batch and BASELINE config 4 (1 M instances over 8 GPUs)
mkdir -p gpurun_out
for w in standing_4096 trot_65536 mixed_terrain_1m; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --workload $w > gpurun_out/s8_$w.json 2> gpurun_out/s8_$w.err
done
python - <<'PY'
import json
for w in ("standing_4096","trot_65536","mixed_terrain_1m"):
    try:
        d=json.loads(open("gpurun_out/s8_%s.json"%w).read().strip().splitlines()[-1])
        print(w, "value %.0f e2e %.0f n_gpus %s ms/step %.3f global_batch %s" % (d["value"], d["e2e"]["value"], d["n_gpus"], d["ms_per_step"], d["config"]["global_batch"]))
    except Exception as e:
        print(w, "ERR", e); print(open("gpurun_out/s8_%s.err"%w).read()[-800:])
PY
