# A/B of variant builds of libwbc_b200.so on ONE box (WBC_B200_LIB selects the library): interleaved runs, both workloads
for rep in 1 2; do
  for v in A B C D; do
    L=$PWD/wbc_quadruped_dob_b200/lib/variants/libwbc_b200_$v.so
    [ -f $L ] || continue
    a=$(WBC_B200_LIB=$L timeout 200 python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f' % d['value'])")
    b=$(WBC_B200_LIB=$L timeout 200 python bench.py --workload trot_65536 --steps 8 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f' % d['value'])")
    echo "rep $rep variant $v: 4096 $a   65536 $b"
  done
done
