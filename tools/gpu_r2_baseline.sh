# round-2 baseline of the HEAD kernel on today's box: tests, bit-exact compare against the round-1 dump, bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_env.txt
bash tools/gpu_quick.sh > gpurun_out/r2a_quick.txt 2>&1
python tools/gpu_dump.py compare tools/_exact/r01_ref.npz > gpurun_out/r2a_compare.txt 2>&1
python tools/gpu_dist.py > gpurun_out/r2a_dist.txt 2>&1; python tools/gpu_dist.py trot_65536 >> gpurun_out/r2a_dist.txt 2>&1
cat gpurun_out/r2a_quick.txt gpurun_out/r2a_compare.txt gpurun_out/r2a_dist.txt | tail -40
