# Q-split staged kernel: QQP-only tasks on the Q SMs
mkdir -p gpurun_out
export WBC_SOLVER=staged
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2k_compare.txt 2>&1; tail -3 gpurun_out/r2k_compare.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python tools/gpu_sanitize.py 96 > gpurun_out/r2k_memcheck.log 2>&1; tail -3 gpurun_out/r2k_memcheck.log
{
for nd in 2/5 1/3 1/2; do
  echo "== M share $nd"
  WBC_STAGE_M_NUM_DEN=$nd timeout 200 python tools/gpu_stage_prof.py standing_4096
  WBC_STAGE_M_NUM_DEN=$nd timeout 300 python tools/gpu_stage_prof.py trot_65536
done
echo "== M share 2/5, groups of 2 SMs"
WBC_STAGE_M_GROUP=2 timeout 300 python tools/gpu_stage_prof.py trot_65536
echo "== M share 2/5, slots 2.0"
WBC_STAGE_SLOTS_PER_WARP=2.0 timeout 300 python tools/gpu_stage_prof.py trot_65536
echo "== no roles"
WBC_STAGE_M_NUM_DEN=0/5 timeout 300 python tools/gpu_stage_prof.py trot_65536
} > gpurun_out/r2k_stage_prof.txt 2>&1
cat gpurun_out/r2k_stage_prof.txt
cp wbc_quadruped_dob_b200/lib/libwbc_b200.so gpurun_out/r2k_lib.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_solve_staged_kernel" -s 3 -c 1 -o gpurun_out/r2k_staged -f python bench.py --workload trot_65536 --per-gpu 16384 --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2k_ncu.log 2>&1; tail -2 gpurun_out/r2k_ncu.log
