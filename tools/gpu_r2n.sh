mkdir -p gpurun_out
export WBC_SOLVER=staged
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2n_compare.txt 2>&1; tail -2 gpurun_out/r2n_compare.txt
{
for roles in 0,2/5 3,7/25; do
  echo "== roles (nS,nP/den) $roles, groups of 2"
  WBC_STAGE_ROLES=$roles timeout 300 python tools/gpu_stage_prof.py trot_65536
done
} > gpurun_out/r2n_roles.txt 2>&1
grep -E "^==|solve kernel|sum of task|mean task" gpurun_out/r2n_roles.txt
cp wbc_quadruped_dob_b200/lib/libwbc_b200.so gpurun_out/r2n_lib.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wbc_solve_staged_kernel" -s 3 -c 1 -o gpurun_out/r2n_staged -f python bench.py --workload trot_65536 --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2n_ncu.log 2>&1; tail -2 gpurun_out/r2n_ncu.log
