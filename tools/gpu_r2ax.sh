# second round of inlining variants on top of the adopted one (U2 = the shipped library; U = the same code built with the old macro set):
# X = + eval4, quadratic_model, explore, step_and_move inlined (2-3 call sites each), Y = + qqp_optimize_fast inlined (one call site per kernel),
# Z = + symv inlined (5 call sites).  S = the library before any of it.
for v in ${VARIANTS:-U2 X Y Z S U2 X Y Z}; do
  L=$PWD/wbc_quadruped_dob_b200/lib/variants/libwbc_b200_$v.so
  a=$(WBC_B200_LIB=$L timeout 200 python bench.py --no-cpu-baseline --no-also --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms)' % (d['value'], d['roofline']['kernel_ms']))")
  b=$(WBC_B200_LIB=$L timeout 200 python bench.py --workload trot_65536 --steps 8 --no-cpu-baseline --no-also 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms)' % (d['value'], d['roofline']['kernel_ms']))")
  echo "variant $v: 4096 $a   65536 $b"
done | tee gpurun_out/r2ax_ab.txt
for v in U2 ${CMPV:-X Y Z}; do WBC_B200_LIB=$PWD/wbc_quadruped_dob_b200/lib/variants/libwbc_b200_$v.so timeout 300 python tools/gpu_dump.py compare tools/_exact/r02_ref.npz 2>&1 | tail -1 | sed "s/^/bit-exactness variant $v: /"; done | tee -a gpurun_out/r2ax_ab.txt
