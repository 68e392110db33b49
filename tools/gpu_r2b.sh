# round 2, step b: shared-memory diet (12 resident solver warps per SM).  Bit-exact compare first, then tests and bench.
mkdir -p gpurun_out
python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2b_compare.txt 2>&1
tail -3 gpurun_out/r2b_compare.txt
python - <<'PY' > gpurun_out/r2b_shape.txt 2>&1
import sys; sys.path.insert(0, ".")
from wbc_quadruped_dob_b200 import api
b = api.WbcBatch(max_batch=4096, device=0); print("solver shape (ctas/SM, smem, grid):", b.solver_shape()); b.close()
PY
cat gpurun_out/r2b_shape.txt
if true; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py 96 > gpurun_out/r2b_memcheck.log 2>&1; tail -30 gpurun_out/r2b_memcheck.log
fi
bash tools/gpu_quick.sh > gpurun_out/r2b_quick.txt 2>&1; cat gpurun_out/r2b_quick.txt
for k in 8 10 11; do
  WBC_SOLVE_CTAS_PER_SM=$k timeout 300 python bench.py --steps 20 --no-cpu-baseline 2>> gpurun_out/occ.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas/SM $k  4096: value %.0f solve_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))"
  WBC_SOLVE_CTAS_PER_SM=$k timeout 300 python bench.py --workload trot_65536 --steps 5 --no-cpu-baseline 2>> gpurun_out/occ.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas/SM $k  65536: value %.0f solve_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))"
done > gpurun_out/r2b_occ.txt 2>&1
cat gpurun_out/r2b_occ.txt
