mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bf16_engine.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r2_6_pytest.txt
timeout 600 python tools/conv_ablate.py 2>&1 | tee gpurun_out/r2_6_ablate.txt | tail -8
