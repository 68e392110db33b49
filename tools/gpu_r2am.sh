# A/B of variant builds, interleaved on one box (WBC_B200_LIB selects the library): VARIANTS names the libraries under lib/variants/, CMPVARIANTS the ones
# whose outputs are compared bit for bit with the stored dump, OUT the result file.  Defaults = the first use: compile flags on top of -maxrregcount=128 (A),
# G = -extra-device-vectorization, H = -Xptxas -dlcm=cg, I = -restrict.
for rep in 1 2; do
  for v in ${VARIANTS:-A G H I}; do
    L=$PWD/wbc_quadruped_dob_b200/lib/variants/libwbc_b200_$v.so
    a=$(WBC_B200_LIB=$L timeout 200 python bench.py --no-cpu-baseline --no-also --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms, front %.4f ms)' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['front_kernel_ms']))")
    b=$(WBC_B200_LIB=$L timeout 200 python bench.py --workload trot_65536 --steps 8 --no-cpu-baseline --no-also 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms, front %.4f ms)' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['front_kernel_ms']))")
    echo "rep $rep variant $v: 4096 $a   65536 $b"
  done
done | tee gpurun_out/${OUT:-r2am_ab}.txt
for v in ${CMPVARIANTS:-G H I}; do
  WBC_B200_LIB=$PWD/wbc_quadruped_dob_b200/lib/variants/libwbc_b200_$v.so timeout 300 python tools/gpu_dump.py compare tools/_exact/r02_ref.npz 2>&1 | tail -1 | sed "s/^/bit-exactness variant $v: /"
done | tee -a gpurun_out/${OUT:-r2am_ab}.txt
