# compute-sanitizer on the FINAL library (-maxrregcount=128 changed the register allocation and the spills): memcheck and racecheck, both solver kernels
mkdir -p gpurun_out
for k in mono staged; do
  WBC_SOLVER=$k timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python tools/gpu_sanitize.py 96 > gpurun_out/r2aq_memcheck_$k.log 2>&1; echo "memcheck $k:"; tail -2 gpurun_out/r2aq_memcheck_$k.log
  WBC_SOLVER=$k timeout 400 compute-sanitizer --tool racecheck --print-limit 10 python tools/gpu_sanitize.py 48 > gpurun_out/r2aq_racecheck_$k.log 2>&1; echo "racecheck $k:"; tail -2 gpurun_out/r2aq_racecheck_$k.log
done
