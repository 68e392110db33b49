# round 2, step y: racecheck after the step_and_move fix (both solver kernels); bit-exactness of the solver against the round-1 dump
# through the thread-per-instance front kernel (whose arithmetic has not changed); a new dump with the default (leg-parallel) front
mkdir -p gpurun_out
WBC_FRONT=thread timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz 2>&1 | head -4 | tee gpurun_out/r2y_compare_thread.txt
timeout 300 python tools/gpu_dump.py dump gpurun_out/r02_ref.npz 2>&1 | tail -1
for k in mono staged; do
  WBC_SOLVER=$k timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 python tools/gpu_sanitize.py 48 > gpurun_out/r2y_racecheck_$k.log 2>&1; echo "racecheck $k:"; tail -2 gpurun_out/r2y_racecheck_$k.log
done
timeout 300 python bench.py --workload standing_4096 --steps 100 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/r2y_x.json 2>> gpurun_out/r2y_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_x.json").read().strip().splitlines()[-1])
print("standing_4096 value %.0f e2e %.0f solve_ms %.3f front_ms %.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["front_kernel_ms"]))
PY
