# round 2, step x: the round's evidence run on one GPU: tests, default bench (headline + also-lines + cpu baseline), reference arm,
# rollout and sweep workloads, single robot, launch list, ncu --set full of the three hot kernels
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -30) > gpurun_out/r2x_pytest.log; tail -3 gpurun_out/r2x_pytest.log
timeout 300 python tools/gpu_dump.py compare tools/_exact/r02_ref.npz > gpurun_out/r2x_compare.txt 2>&1; tail -1 gpurun_out/r2x_compare.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench.err; tail -c 300 gpurun_out/r2x_bench_default.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_ref.json 2>> gpurun_out/r2x_bench.err; tail -c 400 gpurun_out/r2x_bench_ref.json; echo
timeout 600 python bench.py --workload trot_rollout --steps 30 --no-cpu-baseline > gpurun_out/r2x_bench_rollout_4096.json 2>> gpurun_out/r2x_bench.err
timeout 600 python bench.py --workload trot_rollout --per-gpu 65536 --steps 8 --no-cpu-baseline > gpurun_out/r2x_bench_rollout_65536.json 2>> gpurun_out/r2x_bench.err
timeout 900 python bench.py --workload push_sweep --per-gpu 32768 --sweep-cycles 400 > gpurun_out/r2x_bench_sweep_32768.json 2>> gpurun_out/r2x_bench.err
timeout 300 python bench.py --workload trot_replay_single > gpurun_out/r2x_bench_single.json 2>> gpurun_out/r2x_bench.err
python - <<'PY'
import json
for f in ("rollout_4096","rollout_65536","sweep_32768","single"):
    try:
        d=json.loads(open("gpurun_out/r2x_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms/step %.4f fifo %s e2e %.0f" % (d["value"], d["ms_per_step"], d.get("value_fifo"), d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r2x_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_(front_leg|solve)_kernel" -s 6 -c 2 -o gpurun_out/r2x_full_4096 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2x_ncu_full.log 2>&1; tail -1 gpurun_out/r2x_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_(front_leg|solve_staged)_kernel" -s 6 -c 2 -o gpurun_out/r2x_full_65536 -f python bench.py --workload trot_65536 --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2x_ncu_full2.log 2>&1; tail -1 gpurun_out/r2x_ncu_full2.log
cp wbc_quadruped_dob_b200/lib/libwbc_b200.so gpurun_out/r2x_lib.so
