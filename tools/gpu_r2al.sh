# A/B: -maxrregcount=128 (variant B) against the shipped build (A), interleaved on one box
for rep in 1 2; do
  for v in A B C D E F; do
    L=$PWD/wbc_quadruped_dob_b200/lib/variants/libwbc_b200_$v.so
    a=$(WBC_B200_LIB=$L timeout 200 python bench.py --no-cpu-baseline --no-also --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms)' % (d['value'], d['roofline']['kernel_ms']))")
    b=$(WBC_B200_LIB=$L timeout 200 python bench.py --workload trot_65536 --steps 8 --no-cpu-baseline --no-also 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms)' % (d['value'], d['roofline']['kernel_ms']))")
    echo "rep $rep variant $v: 4096 $a   65536 $b"
  done
done | tee gpurun_out/r2al_ab.txt
