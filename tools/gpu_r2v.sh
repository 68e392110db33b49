# round 2, step v: leg-parallel front kernel after the register/reciprocal changes: full GPU suite, timing, ncu of the front kernel,
# and a new bit-exactness reference dump (the front kernel's rounding changed; the solver's did not: staged == mono stays a test)
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/r2v_pytest.log; tail -4 gpurun_out/r2v_pytest.log
timeout 300 python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2v_compare_old.txt 2>&1; head -8 gpurun_out/r2v_compare_old.txt
timeout 300 python tools/gpu_dump.py compare tools/_exact/r02_ref.npz | head -3
for wl in standing_4096 trot_65536; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2v_x.json 2>> gpurun_out/r2v_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2v_x.json").read().strip().splitlines()[-1])
print("$wl value %.0f e2e %.0f solve_ms %.3f front_ms %.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["front_kernel_ms"]))
PY
done
timeout 300 python bench.py --workload trot_replay_single --no-cpu-baseline > gpurun_out/r2v_single.json 2>> gpurun_out/r2v_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2v_single.json").read().strip().splitlines()[-1])
print("single robot: ms_per_step %.4f p50 %s" % (d["ms_per_step"], d.get("p50_ms")))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wbc_front_leg_kernel" -s 4 -c 1 -o gpurun_out/r2v_front -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2v_ncu.log 2>&1; tail -1 gpurun_out/r2v_ncu.log
