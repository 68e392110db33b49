mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"wbc_solve_kernel" -s 3 -c 1 -o gpurun_out/coh_same -f python tools/gpu_coherence.py same > gpurun_out/coh_same.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"wbc_solve_kernel" -s 3 -c 1 -o gpurun_out/coh_varied -f python tools/gpu_coherence.py varied > gpurun_out/coh_varied.log 2>&1
tail -2 gpurun_out/coh_same.log gpurun_out/coh_varied.log
