# round 2, step ac: express lanes as the default for small batches: exactness, latencies, crossover against stage tasks, bench lines
mkdir -p gpurun_out
timeout 300 python tools/gpu_dump.py compare tools/_exact/r02_ref.npz 2>&1 | head -3 | tee gpurun_out/r2ac_compare.txt
for x in 0 9,4,1; do echo "== WBC_EXPRESS=$x"; WBC_EXPRESS=$x timeout 300 python tools/gpu_tail.py standing_4096; done 2>&1 | tee gpurun_out/r2ac_tail.txt
for n in 2048 3000 8192 12288 16384 24576; do
  for k in "mono 0" "mono 9,4,1" "staged 0"; do
    set -- $k
    WBC_SOLVER=$1 WBC_EXPRESS=$2 timeout 300 python bench.py --workload trot_65536 --per-gpu $n --steps 8 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2ac_x.json 2>> gpurun_out/r2ac_bench.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2ac_x.json").read().strip().splitlines()[-1])
print("trot n $n $1 express=$2 value %.0f solve_ms %.3f" % (d["value"], d["roofline"]["kernel_ms"]))
PY
  done
done | tee gpurun_out/r2ac_crossover.txt
for w in standing_4096 trot_rollout; do
  for x in 0 9,4,1; do
    WBC_EXPRESS=$x timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/r2ac_x.json 2>> gpurun_out/r2ac_bench.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2ac_x.json").read().strip().splitlines()[-1])
print("$w express=$x value %.0f e2e %.0f fifo %s solve_ms %.3f" % (d["value"], d["e2e"]["value"], d.get("value_fifo"), d["roofline"]["kernel_ms"]))
PY
  done
done | tee gpurun_out/r2ac_bench.txt
