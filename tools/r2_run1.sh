# round-2 first GPU pass: parity of the new memory-bound kernels, their isolated rooflines, two-issuer conv A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_1_smi.txt
timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_1_pytest.txt; cat gpurun_out/r2_1_pytest.txt
VXM_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err; tail -c 300 gpurun_out/r2_1_bench.err
VXM_B200_TCS2=1 timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bf16_engine.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2_1_pytest_tcs2.txt; cat gpurun_out/r2_1_pytest_tcs2.txt
VXM_B200_TCS2=1 VXM_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --no-kernels > gpurun_out/r2_1_bench_tcs2.json 2> gpurun_out/r2_1_bench_tcs2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_1_bench.json", "gpurun_out/r2_1_bench_tcs2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f e2e %.1f conv_ms %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_step"]))
        for k, v in d.get("kernels", {}).items():
            print("   %-36s %8.1f us  %.3f" % (k, v["us"], v["frac"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
