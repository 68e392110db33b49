mkdir -p gpurun_out
{ timeout 200 python tools/gpu_stage_prof.py standing_4096; timeout 300 python tools/gpu_stage_prof.py trot_65536; WBC_STAGE_M_PERIOD=0 timeout 300 python tools/gpu_stage_prof.py trot_65536; } > gpurun_out/r2f_stage_prof.txt 2>&1
cat gpurun_out/r2f_stage_prof.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wbc_solve_staged_kernel" -s 3 -c 1 -o gpurun_out/r2f_staged -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log; ls -la gpurun_out/r2f_staged.ncu-rep
