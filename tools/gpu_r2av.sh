# small shared helpers (warp sum / max, sqrt, division) inlined (T, -DWBC_SMALL_INLINE) against the shipped library (S): latency-bound cases
for rep in ${REPS:-1 2}; do
  for v in ${VARIANTS:-S T}; do
    L=$PWD/wbc_quadruped_dob_b200/lib/variants/libwbc_b200_$v.so
    a=$(WBC_B200_LIB=$L timeout 200 python bench.py --no-cpu-baseline --no-also --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms)' % (d['value'], d['roofline']['kernel_ms']))")
    s=$(WBC_B200_LIB=$L timeout 200 python bench.py --workload trot_replay_single --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f ms per cycle (p50 %.4f)' % (d['ms_per_step'], d['p50_ms']))")
    b=$(WBC_B200_LIB=$L timeout 200 python bench.py --workload trot_65536 --steps 8 --no-cpu-baseline --no-also 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f (solve %.3f ms)' % (d['value'], d['roofline']['kernel_ms']))")
    echo "rep $rep variant $v: 4096 $a   single robot $s   65536 $b"
  done
done | tee gpurun_out/${OUT:-r2av_ab}.txt
WBC_B200_LIB=$PWD/wbc_quadruped_dob_b200/lib/variants/libwbc_b200_${CMP:-T}.so timeout 300 python tools/gpu_dump.py compare tools/_exact/r02_ref.npz 2>&1 | tail -1 | sed "s/^/bit-exactness variant ${CMP:-T}: /" | tee -a gpurun_out/${OUT:-r2av_ab}.txt
