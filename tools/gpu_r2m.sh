mkdir -p gpurun_out
export WBC_SOLVER=staged
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2m_compare.txt 2>&1; tail -2 gpurun_out/r2m_compare.txt
{
for roles in 0,2/5 1,1/5 3,7/25 3,8/25 2,7/25 4,6/25 0,10/25 1,3/10; do
  echo "== roles (nS,nP/den) $roles, groups of 2"
  WBC_STAGE_ROLES=$roles timeout 300 python tools/gpu_stage_prof.py trot_65536
done
echo "== roles 3,7/25 standing_4096"
WBC_STAGE_ROLES=3,7/25 timeout 300 python tools/gpu_stage_prof.py standing_4096
} > gpurun_out/r2m_roles3.txt 2>&1
grep -E "^==|solve kernel|sum of task" gpurun_out/r2m_roles3.txt
