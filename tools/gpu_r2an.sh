# L1 against resident warps: WBC_SOLVE_L1_CTAS=k keeps k solver warps per SM with an UNPADDED shared-memory request and a carve-out sized
# for them (the rest of the 256 KB is L1); WBC_SOLVE_CTAS_PER_SM=k is the padded form (L1 stays 28 KB).  -dlcm=cg cost 17 %: L1 matters.
run() {
  a=$(env $1 timeout 200 python bench.py --no-cpu-baseline --no-also --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms; %s)' % (d['value'], d['roofline']['kernel_ms'], d['config']['solver_launch'][:40]))")
  b=$(env $1 timeout 200 python bench.py --workload trot_65536 --steps 8 --no-cpu-baseline --no-also 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms; %s)' % (d['value'], d['roofline']['kernel_ms'], d['config']['solver_launch'][:40]))")
  echo "$1: 4096 $a   65536 $b"
}
for e in WBC_X=0 WBC_SOLVE_L1_CTAS=12 WBC_SOLVE_L1_CTAS=11 WBC_SOLVE_L1_CTAS=10 WBC_SOLVE_L1_CTAS=9 WBC_SOLVE_L1_CTAS=8 WBC_SOLVE_L1_CTAS=6 WBC_SOLVE_CTAS_PER_SM=10 WBC_SOLVE_CTAS_PER_SM=8 WBC_X=0; do run $e; done | tee gpurun_out/r2an_l1.txt
