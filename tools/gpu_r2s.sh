# round 2, step s: the staged solver at small batches (where do its warps' cycles go), the staged/mono crossover, the re-run of the
# closed-loop trot test
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_fdyn_plant.py -m gpu -x -q -s 2>&1 | tail -8) > gpurun_out/r2s_pytest.log; tail -4 gpurun_out/r2s_pytest.log
{
for roles in 0,2/5 0,0/5; do
  echo "== staged at 4096, roles $roles"
  WBC_SOLVER=staged WBC_STAGE_ROLES=$roles timeout 300 python tools/gpu_stage_prof.py standing_4096
done
} > gpurun_out/r2s_stage_4096.txt 2>&1
grep -E "^==|solve kernel|share of|sum of task|mean task" gpurun_out/r2s_stage_4096.txt
for n in 8192 12288 16384 24576 32768; do
  for k in mono staged; do
    WBC_SOLVER=$k timeout 300 python bench.py --workload trot_65536 --per-gpu $n --steps 8 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2s_x.json 2>> gpurun_out/r2s_bench.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2s_x.json").read().strip().splitlines()[-1])
print("trot n $n $k value %.0f solve_ms %.3f" % (d["value"], d["roofline"]["kernel_ms"]))
PY
  done
done | tee gpurun_out/r2s_crossover.txt
