# One-GPU validation + evidence pass (run under gpurun): tests, smoke, bench, ncu launch list, ncu --set full of the hot kernels.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
VXM_B200_CONV_ENGINE=bf16 timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-kernels --no-cpu-baseline --no-graph > gpurun_out/launches_final.log 2>&1
for L in rem0 rem2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tcs_kernel|wgrad2_kernel" -s 3 -c 3 -o gpurun_out/r1_final_$L -f python tools/tc_one.py $L > gpurun_out/ncu_final_$L.log 2>&1
done
ls -la gpurun_out | tail -12
