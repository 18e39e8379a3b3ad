mkdir -p gpurun_out
timeout 60 tools/ubench/mma_rate.bin | tee gpurun_out/r2_20_mma_rate.txt
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bf16_engine.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r2_20_pytest.txt
timeout 600 python tools/conv_layers.py 2>&1 | grep -v "^{" | tee gpurun_out/r2_20_layers.txt
VXM_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --no-kernels --no-parity --no-gpu-eager --no-c4 > gpurun_out/r2_20_bench.json 2> gpurun_out/r2_20_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_20_bench.json").read().strip().splitlines()[-1])
print("bench value %.1f e2e %.1f conv_ms %.3f launches/step %s" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_step"], d["launches_per_step"]))
PY
