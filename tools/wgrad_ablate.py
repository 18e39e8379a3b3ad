"""Ablation of the Toeplitz weight-gradient kernel (VXM_B200_WGRAD_DBG bits: 1 no MMAs, 2 no slab copies, 4 no bias sums).  Profiling aid."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxelmorph_b200 import tc

dev = torch.device("cuda:0")
FULL = (160, 192, 224)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda._sleep(200000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(statistics.median(ts) * 1e3, 1)


for cx, cg in ((16, 16), (32, 16), (16, 32), (32, 32)):
    x = torch.randn((1,) + FULL + (cx,), device=dev).to(torch.bfloat16)
    g = torch.randn((1,) + FULL + (cg,), device=dev).to(torch.bfloat16)
    row = {}
    for dbg, label in ((0, "full"), (1, "no_mma"), (2, "no_copy"), (4, "no_bias"), (3, "skeleton+bias"), (7, "skeleton"), (6, "mma_only")):
        os.environ["VXM_B200_WGRAD_DBG"] = str(dbg)
        row[label] = timeit(lambda: tc.conv_wgrad(x, None, g, cx, cg, 3))
    os.environ.pop("VXM_B200_WGRAD_DBG")
    print("wgrad x%d g%d: %s" % (cx, cg, row), flush=True)
