mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_bf16_engine.py -x -q -m gpu 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["ms_per_step"], d["e2e"]["value"])'
VXM_B200_CONV_ENGINE=bf16 timeout 80 python bench.py --steps 10 --warmup 3 --no-kernels --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$P" default
