mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_3_pytest.txt; cat gpurun_out/r2_3_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5
timeout 1500 python bench.py > gpurun_out/r2_3_bench.json 2> gpurun_out/r2_3_bench.err; tail -c 400 gpurun_out/r2_3_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_3_bench.json").read().strip().splitlines()[-1])
    print("bench value %.1f e2e %.1f conv_ms %.3f launches/step %s frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_step"], d["launches_per_step"], d["roofline"]["frac"]))
    print("parity_check", d.get("parity_check")); print("parity_mode", d.get("parity_mode")); print("gpu_eager", d.get("gpu_eager_baseline")); print("cpu", d.get("cpu_baseline"))
    for k, v in d.get("kernels", {}).items():
        print("   %-36s %8.1f us  %.3f" % (k, v["us"], v["frac"]))
    c4 = d.get("c4_sweep") or {}
    for grp in ("warp", "vecint"):
        for k, v in (c4.get(grp) or {}).items():
            print("   c4 %-28s %s" % (k, v))
except Exception as e:
    print("bench unreadable", e)
PY
for K in ncc9_kernel vecint_fwd_fast warp_fwd_fast warp_bwd_fast resize_bwd_march vecint_bwd_fast; do
  timeout 300 ncu --set full --clock-control none -k regex:$K -s 1 -c 1 -f -o gpurun_out/r2_3_$K python tools/r2_memprof.py launch > gpurun_out/r2_3_ncu_$K.log 2>&1
done
ls -la gpurun_out | tail -12; du -sh gpurun_out
