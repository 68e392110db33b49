mkdir -p gpurun_out
export WBC_SOLVER=staged
{
for g in 2 4 6 8 18; do for nd in 2/5 3/8 1/3; do
  echo "== M share $nd, groups of $g SMs"
  WBC_STAGE_M_GROUP=$g WBC_STAGE_M_NUM_DEN=$nd timeout 300 python tools/gpu_stage_prof.py trot_65536 | head -1
done; done
for nd in 2/5 3/8 1/3; do
  echo "== 4096: M share $nd, groups of 2 SMs"
  WBC_STAGE_M_GROUP=2 WBC_STAGE_M_NUM_DEN=$nd timeout 300 python tools/gpu_stage_prof.py standing_4096 | head -1
done
echo "== mixed 1M shard, 2/5 g2"
WBC_STAGE_M_GROUP=2 timeout 300 python tools/gpu_stage_prof.py mixed_terrain_1m | head -5
} > gpurun_out/r2l_roles.txt 2>&1
cat gpurun_out/r2l_roles.txt
