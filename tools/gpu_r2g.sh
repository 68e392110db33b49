mkdir -p gpurun_out
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2g_compare.txt 2>&1; tail -2 gpurun_out/r2g_compare.txt
{
for spw in 1.1 1.25 1.5 2.0; do
  echo "== slots per warp $spw"
  WBC_STAGE_SLOTS_PER_WARP=$spw timeout 200 python tools/gpu_stage_prof.py standing_4096
  WBC_STAGE_SLOTS_PER_WARP=$spw timeout 300 python tools/gpu_stage_prof.py trot_65536
done
echo "== no roles, 1.25"
WBC_STAGE_M_PERIOD=0 WBC_STAGE_SLOTS_PER_WARP=1.25 timeout 300 python tools/gpu_stage_prof.py trot_65536
} > gpurun_out/r2g_stage_prof.txt 2>&1
cat gpurun_out/r2g_stage_prof.txt
