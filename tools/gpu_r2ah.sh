for x in 9,4,1,2048,12 9,4,1,2048,10 0; do echo "== WBC_EXPRESS=$x"; WBC_EXPRESS=$x timeout 300 python tools/gpu_warp_timeline.py standing_4096; done 2>&1 | tee gpurun_out/r2ah_timeline.txt
