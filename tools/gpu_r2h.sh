# round 2, step h: full validation of the round's state: GPU tests (old + new), sanitizer, default bench with `also`, rollout and sweep benches
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25) > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2h_compare.txt 2>&1; tail -2 gpurun_out/r2h_compare.txt
timeout 600 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench.err; tail -c 2500 gpurun_out/r2h_bench_default.json; tail -3 gpurun_out/r2h_bench.err
timeout 600 python bench.py --workload trot_rollout --steps 30 > gpurun_out/r2h_bench_rollout_4096.json 2>> gpurun_out/r2h_bench.err; tail -c 1500 gpurun_out/r2h_bench_rollout_4096.json
timeout 600 python bench.py --workload trot_rollout --per-gpu 65536 --steps 8 --no-cpu-baseline > gpurun_out/r2h_bench_rollout_65536.json 2>> gpurun_out/r2h_bench.err; tail -c 1500 gpurun_out/r2h_bench_rollout_65536.json
timeout 900 python bench.py --workload push_sweep --per-gpu 32768 --sweep-cycles 400 > gpurun_out/r2h_bench_sweep_32768.json 2>> gpurun_out/r2h_bench.err; tail -c 3000 gpurun_out/r2h_bench_sweep_32768.json
tail -5 gpurun_out/r2h_bench.err
