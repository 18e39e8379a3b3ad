mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tcs|wgrad2_kernel" -c 14 -f -o gpurun_out/r2_conv_full python tools/conv_layers.py --once rem0 rem1 rem2 "fold: enc0" "fold: flow" > gpurun_out/r2_ncu_conv.log 2>&1
tail -2 gpurun_out/r2_ncu_conv.log; ls -la gpurun_out/r2_conv_full.ncu-rep
python tools/ncu_summary.py gpurun_out/r2_conv_full.ncu-rep > gpurun_out/r2_conv_ncu.md 2> gpurun_out/r2_conv_ncu.err; wc -l gpurun_out/r2_conv_ncu.md
ncu -i gpurun_out/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_full_raw.csv 2>/dev/null; ls -la gpurun_out/
[ $(stat -c %s gpurun_out/r2_conv_full.ncu-rep) -gt 50000000 ] && rm gpurun_out/r2_conv_full.ncu-rep
du -sh gpurun_out
