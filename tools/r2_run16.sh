mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_16_pytest.txt; cat gpurun_out/r2_16_pytest.txt
timeout 600 python tools/conv_layers.py 2>&1 | grep "fold\|plain" | tee gpurun_out/r2_16_layers.txt
VXM_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --no-kernels --no-parity --no-gpu-eager --no-c4 > gpurun_out/r2_16_bench.json 2> gpurun_out/r2_16_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_16_bench.json").read().strip().splitlines()[-1])
print("bench value %.1f e2e %.1f conv_ms %.3f launches/step %s" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_step"], d["launches_per_step"]))
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_16_launches.csv python bench.py --steps 2 --warmup 3 --no-kernels --no-cpu-baseline --no-graph --no-parity --no-gpu-eager --no-c4 > gpurun_out/r2_16_launches.log 2>&1
