mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bf16_engine.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r2_9_pytest.txt
timeout 900 python tools/conv_ablate.py 2>&1 | tee gpurun_out/r2_9_ablate.txt | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['layer'])
    cfgs = []
    for k in d['us']:
        c = k.split(':')[0]
        if c not in cfgs: cfgs.append(c)
    for c in cfgs: print('   ', c, {k.split(':')[1]: v for k, v in d['us'].items() if k.startswith(c + ':')})
"
