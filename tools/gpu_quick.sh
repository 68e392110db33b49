mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/q_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/q_bench_4096.json 2> gpurun_out/q_bench.err
timeout 300 python bench.py --workload trot_65536 --steps 10 --no-cpu-baseline > gpurun_out/q_bench_65536.json 2>> gpurun_out/q_bench.err
cat gpurun_out/q_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/q_bench_4096.json","gpurun_out/q_bench_65536.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f solve_ms %.3f front_ms %.3f frac %.4f nchol %.3f fail %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["front_kernel_ms"], d["roofline"]["frac"], d["stats"]["mean_ncholesky"], d["stats"]["solver_failures"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/q_bench.err
timeout 300 python bench.py --fifo --no-cpu-baseline 2>> gpurun_out/q_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fifo 4096 value %.0f solve_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))"
