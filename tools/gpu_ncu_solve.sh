# one ncu --set full capture of the solver kernel (4096 standing instances); the report comes back in gpurun_out/
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_solve_kernel" -s 4 -c 1 -o gpurun_out/full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
