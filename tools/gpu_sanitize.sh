mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py 96 > gpurun_out/san_memcheck.log 2>&1; tail -4 gpurun_out/san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_sanitize.py 48 > gpurun_out/san_racecheck.log 2>&1; tail -4 gpurun_out/san_racecheck.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python tools/gpu_sanitize.py 48 > gpurun_out/san_initcheck.log 2>&1; tail -4 gpurun_out/san_initcheck.log
