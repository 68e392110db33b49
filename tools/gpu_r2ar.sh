# spread of the SM instruction-cache behaviour over the SMs of the stage-task kernel (roles: QQP / POST / SETUP SMs run different code)
M=sm__icc_requests.sum,sm__icc_requests.min,sm__icc_requests.max,sm__icc_requests_lookup_miss.sum,sm__icc_requests_lookup_miss.min,sm__icc_requests_lookup_miss.max,sm__icc_requests_lookup_hit.min,sm__icc_requests_lookup_hit.max,sm__inst_executed.sum,sm__inst_executed.min,sm__inst_executed.max,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_hit.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,sm__issue_active.min,sm__issue_active.max,sm__issue_active.avg,sm__cycles_active.avg
for R in "" "0,0/1"; do
  WBC_STAGE_ROLES=$R timeout 600 ncu --metrics $M --clock-control none -k regex:wbc_solve_staged_kernel -s 6 -c 1 --csv --log-file gpurun_out/r2ar_icc_roles_${R//[,\/]/_}.csv python bench.py --workload trot_65536 --steps 2 --warmup 3 --no-cpu-baseline --no-also > /dev/null 2>&1
done
timeout 600 ncu --metrics $M --clock-control none -k regex:wbc_solve_kernel -s 6 -c 1 --csv --log-file gpurun_out/r2ar_icc_mono4096.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > /dev/null 2>&1
tail -n +1 gpurun_out/r2ar_icc_*.csv | cut -c1-300 | tail -80
