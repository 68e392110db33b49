# round 2, step t: qqp_stats bank-conflict fix (bit-exactness + timing), two against three SM roles at 65536 / 131072, closed-loop trot test
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_fdyn_plant.py -m gpu -x -q -s 2>&1 | tail -6) > gpurun_out/r2t_pytest.log; tail -3 gpurun_out/r2t_pytest.log
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2t_compare.txt 2>&1; tail -1 gpurun_out/r2t_compare.txt
run() { # label, env..., args
  label=$1; shift
  env "$@" > /dev/null 2>&1
}
for rep in 1 2; do
for cfg in "2roles:0,2/5" "3roles:3,7/25" "3roles_b:2,6/20" "3roles_c:1,3/10"; do
  name=${cfg%%:*}; roles=${cfg##*:}
  for wl in trot_65536 mixed_terrain_1m; do
    WBC_STAGE_ROLES=$roles timeout 300 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2t_x.json 2>> gpurun_out/r2t_bench.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2t_x.json").read().strip().splitlines()[-1])
print("$name $roles $wl value %.0f solve_ms %.3f" % (d["value"], d["roofline"]["kernel_ms"]))
PY
  done
done
done | tee gpurun_out/r2t_roles.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/r2t_4096.json 2>> gpurun_out/r2t_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2t_4096.json").read().strip().splitlines()[-1])
print("standing_4096 value %.0f solve_ms %.3f" % (d["value"], d["roofline"]["kernel_ms"]))
PY
