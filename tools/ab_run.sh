mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-kernels --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["ms_per_step"])'
export VXM_B200_CONV_ENGINE=bf16
timeout 80 $B 2>/dev/null | tail -1 | python -c "$P" default
VXM_B200_NCC_ZCHUNK=32 timeout 80 $B 2>/dev/null | tail -1 | python -c "$P" z32
VXM_B200_NCC_ZCHUNK=54 timeout 80 $B 2>/dev/null | tail -1 | python -c "$P" z54
