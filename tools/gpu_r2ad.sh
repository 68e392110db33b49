for x in 9,1,1 9,2,1 5,1,1 5,2,1 9,1,2; do echo "== WBC_EXPRESS=$x"; WBC_EXPRESS=$x timeout 300 python tools/gpu_tail2.py | head -9; done 2>&1 | tee gpurun_out/r2ad_tail3.txt
