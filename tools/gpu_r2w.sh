# round 2, step w: compute-sanitizer on the round's kernels: memcheck and racecheck for both solver kernels, initcheck once
mkdir -p gpurun_out
for k in mono staged; do
  WBC_SOLVER=$k timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/gpu_sanitize.py 96 > gpurun_out/r2w_memcheck_$k.log 2>&1; echo "memcheck $k:"; tail -2 gpurun_out/r2w_memcheck_$k.log
  WBC_SOLVER=$k timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 python tools/gpu_sanitize.py 48 > gpurun_out/r2w_racecheck_$k.log 2>&1; echo "racecheck $k:"; tail -2 gpurun_out/r2w_racecheck_$k.log
done
WBC_FRONT=thread timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/gpu_sanitize.py 48 > gpurun_out/r2w_memcheck_thread.log 2>&1; echo "memcheck front=thread:"; tail -2 gpurun_out/r2w_memcheck_thread.log
timeout 900 compute-sanitizer --tool initcheck --print-limit 10 python tools/gpu_sanitize.py 48 > gpurun_out/r2w_initcheck.log 2>&1; echo "initcheck:"; tail -2 gpurun_out/r2w_initcheck.log
