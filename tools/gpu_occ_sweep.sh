# how does solver throughput scale with resident warps per SM?  (WBC_SOLVE_CTAS_PER_SM pads the shared-memory request)
mkdir -p gpurun_out
for k in 2 4 6 7 8; do
  WBC_SOLVE_CTAS_PER_SM=$k timeout 300 python bench.py --workload trot_65536 --steps 5 --no-cpu-baseline 2>> gpurun_out/occ.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas/SM $k  65536: value %.0f solve_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))"
done
