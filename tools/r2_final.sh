# round-2 validation + evidence pass (1x B200): full GPU suite, smoke, full bench line, ncu launch list with DRAM bytes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_f_smi.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_f_pytest.txt; cat gpurun_out/r2_f_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4 | tee gpurun_out/r2_f_smoke.txt
VXM_BENCH_VERBOSE=1 timeout 1500 python bench.py > gpurun_out/r2_f_bench.json 2> gpurun_out/r2_f_bench.err; tail -c 300 gpurun_out/r2_f_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_f_bench.json").read().strip().splitlines()[-1])
    print("bench value %.1f e2e %.1f conv_ms %.3f launches/step %s frac %.3f (sustained %.3f)" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_step"], d["launches_per_step"], d["roofline"]["frac"], d["roofline"]["frac_of_sustained"]))
    print("parity_check", d.get("parity_check")); print("parity_mode", d.get("parity_mode")); print("gpu_eager", d.get("gpu_eager_baseline")); print("cpu", d.get("cpu_baseline")); print("clocks", d.get("clocks"))
    for k, v in d.get("kernels", {}).items():
        print("   %-36s %8.1f us  %.3f" % (k, v["us"], v["frac"]))
except Exception as e:
    print("bench unreadable", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_f_launches.csv python bench.py --steps 2 --warmup 3 --no-kernels --no-cpu-baseline --no-graph --no-parity --no-gpu-eager --no-c4 > gpurun_out/r2_f_launches.log 2>&1
timeout 300 python tools/conv_layers.py 2>&1 | grep -v "^{" > gpurun_out/r2_f_layers.txt
