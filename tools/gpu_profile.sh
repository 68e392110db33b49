# full GPU check: tests, three bench workloads, ncu launch list + one --set full capture of both kernels
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/q_pytest.log
timeout 300 python bench.py > gpurun_out/q_bench_4096.json 2> gpurun_out/q_bench.err
timeout 300 python bench.py --workload trot_65536 --steps 10 --no-cpu-baseline > gpurun_out/q_bench_65536.json 2>> gpurun_out/q_bench.err
timeout 300 python bench.py --workload mixed_terrain_1m --steps 10 --no-cpu-baseline > gpurun_out/q_bench_1m.json 2>> gpurun_out/q_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_(front|solve)_kernel" -s 6 -c 2 -o gpurun_out/full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/q_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/q_bench_4096.json","gpurun_out/q_bench_65536.json","gpurun_out/q_bench_1m.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f solve_ms %.3f front_ms %.3f frac %.4f nchol %.3f fail %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["front_kernel_ms"], d["roofline"]["frac"], d["stats"]["mean_ncholesky"], d["stats"]["solver_failures"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/q_bench.err
