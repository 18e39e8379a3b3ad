mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r2_7_pytest.txt
timeout 300 python tools/r2_memprof.py time > gpurun_out/r2_7_memtime.json 2> gpurun_out/r2_7_memtime.txt; cat gpurun_out/r2_7_memtime.txt
for K in vecint_fwd_fast vecint_bwd_fast ncc9_kernel warp_bwd_fast warp_fwd_fast resize_bwd_colsm; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/r2_7_$K python tools/r2_memprof.py launch > gpurun_out/r2_7_ncu_$K.log 2>&1
done
ls -la gpurun_out | tail -12
