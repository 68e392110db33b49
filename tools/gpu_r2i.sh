mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25) > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
timeout 600 python bench.py --no-also > gpurun_out/r2i_bench_4096.json 2> gpurun_out/r2i_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2i_bench_4096.json").read().strip().splitlines()[-1])
print("4096 value %.0f e2e %.0f kernel_ms %.3f %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["solver_launch"]))
PY
timeout 900 python bench.py --workload push_sweep --per-gpu 32768 --sweep-cycles 400 > gpurun_out/r2i_bench_sweep_32768.json 2>> gpurun_out/r2i_bench.err; tail -c 600 gpurun_out/r2i_bench_sweep_32768.json
tail -5 gpurun_out/r2i_bench.err
