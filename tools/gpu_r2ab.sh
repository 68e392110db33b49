# round 2, step ab: express lanes for the one-warp-per-solve kernel at 4096 instances: a sweep of the settings (flop-count ranking)
mkdir -p gpurun_out
for x in 0 9,4,1 9,4,1.5 9,4,0.75 7,4,1 6,4,1 5,4,1 4,4,1 9,6,1 9,3,1 9,5,1 12,4,1 6,6,1 6,3,1; do
  echo "== WBC_EXPRESS=$x"
  WBC_EXPRESS=$x timeout 300 python tools/gpu_tail.py standing_4096 | head -1
done 2>&1 | tee gpurun_out/r2ab_express2.txt
