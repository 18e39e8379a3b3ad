mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bf16_engine.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r2_8_pytest.txt
timeout 600 python tools/conv_ablate.py 2>&1 | tee gpurun_out/r2_8_ablate.txt | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    u = d['us']; print(d['layer'], {k[5:]: v for k, v in u.items() if k.startswith('tma1_')})
"
