# round 2, step r: new GPU tests (observer forms, foot-wrench map, closed-loop trot through the dynamics plant, C++ host layer) and
# a sweep of the staged solver's slots in flight per warp (L2 residency of the hand-over state) at 65536 instances
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_observer_ext.py tests/test_fdyn_plant.py tests/test_cpp_host_layer.py tests/test_capi_symbols.py -m gpu -x -q -s 2>&1 | tail -25) > gpurun_out/r2r_pytest.log; tail -8 gpurun_out/r2r_pytest.log
for spw in 1.0 1.25 1.5 2.0 3.0; do
  WBC_STAGE_SLOTS_PER_WARP=$spw timeout 300 python bench.py --workload trot_65536 --steps 8 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2r_slots_$spw.json 2>> gpurun_out/r2r_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2r_slots_$spw.json").read().strip().splitlines()[-1])
print("slots/warp $spw value %.0f solve_ms %.3f" % (d["value"], d["roofline"]["kernel_ms"]))
PY
done
