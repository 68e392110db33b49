# round 2, step u: the leg-parallel front kernel: equivalence test, full GPU suite, timing at 4096 / 65536 / single robot
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -s -k "leg_parallel" 2>&1 | tail -15) > gpurun_out/r2u_front.log; tail -6 gpurun_out/r2u_front.log
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2u_pytest.log; tail -4 gpurun_out/r2u_pytest.log
for k in leg thread; do
  for wl in standing_4096 trot_65536; do
    WBC_FRONT=$k timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2u_x.json 2>> gpurun_out/r2u_bench.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2u_x.json").read().strip().splitlines()[-1])
print("front=$k $wl value %.0f e2e %.0f solve_ms %.3f front_ms %.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["front_kernel_ms"]))
PY
  done
  WBC_FRONT=$k timeout 300 python bench.py --workload trot_replay_single --no-cpu-baseline > gpurun_out/r2u_single_$k.json 2>> gpurun_out/r2u_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2u_single_$k.json").read().strip().splitlines()[-1])
print("front=$k single robot: ms_per_step %.4f p50 %s e2e %s" % (d["ms_per_step"], d.get("p50_ms"), d["e2e"]))
PY
done | tee gpurun_out/r2u_front_ab.txt
