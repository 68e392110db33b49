# round 2, step aa: the tail of a 4096-instance step for 8..12 resident solver warps per SM; the node-glue test
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_ros_adapter.py -m gpu -q 2>&1 | tail -3) | tee gpurun_out/r2aa_pytest.log
for w in 6 8 9 10 12; do
  echo "== WBC_SOLVE_CTAS_PER_SM=$w"
  WBC_SOLVE_CTAS_PER_SM=$w timeout 300 python tools/gpu_tail.py standing_4096
done 2>&1 | tee gpurun_out/r2aa_tail.txt
