mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_14_pytest.txt; cat gpurun_out/r2_14_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4 | tee gpurun_out/r2_14_smoke.txt
VXM_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --no-parity --no-gpu-eager --no-c4 > gpurun_out/r2_14_bench.json 2> gpurun_out/r2_14_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_14_bench.json").read().strip().splitlines()[-1])
print("bench value %.1f e2e %.1f conv_ms %.3f launches/step %s frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_step"], d["launches_per_step"], d["roofline"]["frac"]))
for k, v in d.get("kernels", {}).items():
    print("   %-36s %8.1f us  %.3f" % (k, v["us"], v["frac"]))
PY
