import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxelmorph_b200 import tc
dev = torch.device("cuda:0")
def run(name, shape, Ca, Cb, up, Cout):
    D, H, W = shape
    ash = (D // 2, H // 2, W // 2) if up else shape
    xa = torch.randn((1,) + ash + (Ca,), device=dev).to(torch.bfloat16)
    xb = torch.randn((1,) + shape + (Cb,), device=dev).to(torch.bfloat16) if Cb else None
    w = torch.randn((Cout, Ca + Cb, 3, 3, 3), device=dev) * 0.05
    wpk, NP = tc.pack_weights(w); b = torch.zeros(Cout, device=dev)
    ts = []
    for i in range(6):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000); a0.record(); tc.conv_fwd(xa, xb, wpk, NP, b, Cout, 3, up=up, slope=0.2); a1.record(); torch.cuda.synchronize()
        ts.append(a0.elapsed_time(a1))
    print(name, "interleave", os.environ.get("VXM_TC_DBG_INTERLEAVE"), "ms", round(statistics.median(ts[2:]), 3), flush=True)
FULL = (160, 192, 224)
run("rem0", FULL, 32, 16, True, 32); run("rem1", FULL, 32, 0, False, 16); run("rem2", FULL, 16, 0, False, 16)
