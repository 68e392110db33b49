mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25) > gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_pytest.log
b() { lbl=$1; shift; out=$(env "$@" 2>>gpurun_out/r2j_bench.err) || true
  echo "$out" | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lbl value %.0f e2e %.0f solve_ms %.3f front_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['front_kernel_ms']))
except Exception as e: print('$lbl ERR', e)"; }
{
for rep in 1 2; do
b "adaptive 4096" timeout 200 python bench.py --no-cpu-baseline --no-also --steps 30
b "forced12 4096" WBC_SOLVE_CTAS_PER_SM=12 timeout 200 python bench.py --no-cpu-baseline --no-also --steps 30
b "forced8  4096" WBC_SOLVE_CTAS_PER_SM=8 timeout 200 python bench.py --no-cpu-baseline --no-also --steps 30
b "forced10 4096" WBC_SOLVE_CTAS_PER_SM=10 timeout 200 python bench.py --no-cpu-baseline --no-also --steps 30
done
} > gpurun_out/r2j_occ_ab.txt 2>&1; cat gpurun_out/r2j_occ_ab.txt
cp wbc_quadruped_dob_b200/lib/libwbc_b200.so gpurun_out/r2j_lib.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_solve_kernel" -s 4 -c 1 -o gpurun_out/r2j_mono -f env WBC_SOLVE_CTAS_PER_SM=12 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2j_ncu.log 2>&1; tail -2 gpurun_out/r2j_ncu.log
