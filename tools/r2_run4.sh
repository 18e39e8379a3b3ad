mkdir -p gpurun_out
timeout 300 python tools/r2_memprof.py time > gpurun_out/r2_4_memtime.json 2> gpurun_out/r2_4_memtime.txt; cat gpurun_out/r2_4_memtime.txt
VXM_B200_RESIZE_BWD=march timeout 100 python tools/r2_memprof.py time 2>&1 | grep resize_up_bwd
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_bf16_engine.py tests/test_gpu_configs.py -q -m gpu -x 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5
VXM_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --no-kernels --no-parity --no-gpu-eager --no-c4 > gpurun_out/r2_4_bench.json 2> gpurun_out/r2_4_bench.err
VXM_B200_WGRAD_DEFER=0 timeout 600 python bench.py --no-cpu-baseline --no-kernels --no-parity --no-gpu-eager --no-c4 > gpurun_out/r2_4_bench_nodefer.json 2> gpurun_out/r2_4_bench_nodefer.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_4_bench.json", "gpurun_out/r2_4_bench_nodefer.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f e2e %.1f conv_ms %.3f launches/step %s" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_step"], d["launches_per_step"]))
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
timeout 300 ncu --set full --clock-control none -k regex:ncc9_kernel -s 3 -c 1 -f -o gpurun_out/r2_4_ncc9 python tools/r2_memprof.py launch > gpurun_out/r2_4_ncu_ncc9.log 2>&1; ls -la gpurun_out/r2_4_ncc9.ncu-rep
