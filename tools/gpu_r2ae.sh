# round 2, step ae: qqp_stats with every lane on the same row (the 16-way conflict really gone?): exactness, timing, conflict counters
mkdir -p gpurun_out
timeout 300 python tools/gpu_dump.py compare tools/_exact/r02_ref.npz 2>&1 | head -3 | tee gpurun_out/r2ae_compare.txt
for rep in 1 2; do for wl in standing_4096 trot_65536; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/r2ae_x.json 2>> gpurun_out/r2ae_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2ae_x.json").read().strip().splitlines()[-1])
print("$wl value %.0f e2e %.0f solve_ms %.3f front_ms %.4f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["front_kernel_ms"], d["ms_per_step"]))
PY
done; done | tee gpurun_out/r2ae_bench.txt
timeout 600 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum --clock-control none -k regex:"wbc_solve_kernel" -s 3 -c 1 --csv --log-file gpurun_out/r2ae_conflicts.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > /dev/null 2>&1; grep -E "wbc_solve" gpurun_out/r2ae_conflicts.csv | awk -F'","' '{print $(NF-2), $NF}'
