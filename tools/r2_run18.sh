mkdir -p gpurun_out
timeout 60 tools/ubench/mma_rate.bin | tee gpurun_out/r2_18_mma_rate.txt
bash tools/r2_run17.sh
