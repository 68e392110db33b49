# round 2, step e: staged solver kernel (stage tasks, SM roles).  Every command under its own timeout: a hung persistent kernel must not hang the box.
mkdir -p gpurun_out
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2e_compare.txt 2>&1; tail -4 gpurun_out/r2e_compare.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python tools/gpu_sanitize.py 96 > gpurun_out/r2e_memcheck.log 2>&1; tail -3 gpurun_out/r2e_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/gpu_sanitize.py 48 > gpurun_out/r2e_racecheck.log 2>&1; tail -3 gpurun_out/r2e_racecheck.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_pytest.log
b() { # label, env..., args
  lbl=$1; shift
  out=$(env "$@" 2>>gpurun_out/r2e_bench.err) || true
  echo "$out" | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl value %.0f e2e %.0f solve_ms %.3f front_ms %.3f frac %.4f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['front_kernel_ms'], d['roofline']['frac']))
except Exception as e: print('$lbl ERR', e)"
}
{
b "mono    4096 " WBC_SOLVER=mono timeout 200 python bench.py --no-cpu-baseline --steps 20
b "mono    65536" WBC_SOLVER=mono timeout 200 python bench.py --workload trot_65536 --steps 6 --no-cpu-baseline
b "staged3 4096 " timeout 200 python bench.py --no-cpu-baseline --steps 20
b "staged3 65536" timeout 200 python bench.py --workload trot_65536 --steps 6 --no-cpu-baseline
b "staged0 65536" WBC_STAGE_M_PERIOD=0 timeout 200 python bench.py --workload trot_65536 --steps 6 --no-cpu-baseline
b "staged2 65536" WBC_STAGE_M_PERIOD=2 timeout 200 python bench.py --workload trot_65536 --steps 6 --no-cpu-baseline
b "staged4 65536" WBC_STAGE_M_PERIOD=4 timeout 200 python bench.py --workload trot_65536 --steps 6 --no-cpu-baseline
b "staged3g2 65536" WBC_STAGE_M_GROUP=2 timeout 200 python bench.py --workload trot_65536 --steps 6 --no-cpu-baseline
b "staged3g8 65536" WBC_STAGE_M_GROUP=8 timeout 200 python bench.py --workload trot_65536 --steps 6 --no-cpu-baseline
b "staged3 slots4 65536" WBC_STAGE_SLOTS_PER_WARP=4 timeout 200 python bench.py --workload trot_65536 --steps 6 --no-cpu-baseline
b "staged3 1M/8  " timeout 300 python bench.py --workload mixed_terrain_1m --steps 4 --no-cpu-baseline
} > gpurun_out/r2e_bench.txt 2>&1
cat gpurun_out/r2e_bench.txt; tail -5 gpurun_out/r2e_bench.err
