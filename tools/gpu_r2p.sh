# round 2, step p: state after "solver kernel by batch size": GPU tests, bit-exact dump, default bench (with also-lines), launch list,
# ncu --set full of the front kernel + one-warp-per-solve kernel (4096) and of the staged kernel (65536)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25) > gpurun_out/r2p_pytest.log; tail -6 gpurun_out/r2p_pytest.log
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2p_compare.txt 2>&1; tail -2 gpurun_out/r2p_compare.txt
timeout 900 python bench.py > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench.err; tail -c 1500 gpurun_out/r2p_bench_default.json; tail -3 gpurun_out/r2p_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2p_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_(front|solve)_kernel" -s 6 -c 2 -o gpurun_out/r2p_full_4096 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2p_ncu_full.log 2>&1; tail -2 gpurun_out/r2p_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_(front|solve_staged)_kernel" -s 6 -c 2 -o gpurun_out/r2p_full_65536 -f python bench.py --workload trot_65536 --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2p_ncu_full2.log 2>&1; tail -2 gpurun_out/r2p_ncu_full2.log
cp wbc_quadruped_dob_b200/lib/libwbc_b200.so gpurun_out/r2p_lib.so
