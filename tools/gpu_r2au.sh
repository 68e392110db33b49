# SM roles of the stage-task kernel re-checked on the final library (nS,nP/den of SM groups: SETUP, POST, the rest QQP), trot_65536, 8 steps each
for R in 3,7/25 2,7/25 3,8/25 3,6/25 4,7/25 2,8/25 3,7/25; do
  b=$(WBC_STAGE_ROLES=$R timeout 200 python bench.py --workload trot_65536 --steps 8 --no-cpu-baseline --no-also 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms)' % (d['value'], d['roofline']['kernel_ms']))")
  echo "WBC_STAGE_ROLES=$R: $b"
done | tee gpurun_out/r2au_roles.txt
