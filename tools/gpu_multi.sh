# 2-GPU weak-scaling run, the reference arm, and the closed-loop sweep: every bench path on the box
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/m_bench_2gpu.json 2> gpurun_out/m_bench_2gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench_ref.err
timeout 600 python bench.py --workload push_sweep --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_sweep.json 2> gpurun_out/m_bench_sweep.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/m_bench_1gpu.json 2> gpurun_out/m_bench_1gpu.err
python - <<'PY'
import json
for f in ("m_bench_2gpu","m_bench_ref","m_bench_sweep","m_bench_1gpu"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f n_gpus %s ms/step %.3f launches %s cpu %s" % (d["value"], d["e2e"]["value"], d["n_gpus"], d["ms_per_step"], d.get("gpu_launches"), (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-600:])
PY
