# does occupancy help once the instruction cache is coherent?  (all instances copies of one: every warp of an SM in the same phase)
mkdir -p gpurun_out
for k in 6 8 10 12; do
  echo "== resident solver warps per SM: $k"
  WBC_SOLVE_CTAS_PER_SM=$k timeout 300 python tools/gpu_coherence.py all
done > gpurun_out/r2c_coherence_occ.txt 2>&1
cat gpurun_out/r2c_coherence_occ.txt
