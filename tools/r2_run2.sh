mkdir -p gpurun_out
timeout 300 python tools/r2_memprof.py time > gpurun_out/r2_2_memtime.json 2> gpurun_out/r2_2_memtime.txt; cat gpurun_out/r2_2_memtime.txt
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_bf16_engine.py tests/test_shim.py -q -m gpu -s 2>&1 | grep -E "^\[|bf16x3|full size|passed|failed|FAILED|Error|assert" | head -60
VXM_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --no-kernels > gpurun_out/r2_2_bench.json 2> gpurun_out/r2_2_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_2_bench.json").read().strip().splitlines()[-1])
print("bench value %.1f e2e %.1f conv_ms %.3f launches/step %s" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_step"], d["launches_per_step"]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_fwd_fast|warp_bwd_fast|vecint_fwd_fast|vecint_bwd_fast|resize_fwd_kernel|resize_bwd_march|ncc9_kernel" -c 48 -f -o gpurun_out/r2_2_mem python tools/r2_memprof.py launch > gpurun_out/r2_2_ncu.log 2>&1; tail -3 gpurun_out/r2_2_ncu.log; ls -la gpurun_out/r2_2_mem.ncu-rep
