"""Critical-path ablation of the kw-stacked tcgen05 convolution (profiling aid, not a product path).
Times full-resolution layers with parts of the kernel switched off (VXM_B200_TCS_DBG bits: 1 no MMAs, 2 no TMEM read-out,
4 no global stores / mask loads, 8 no slab copies) and with the A operand staged by TMA tensor copies vs cp.async."""
import sys, os, statistics, json
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
    return statistics.median(ts)


def layer(name, shape, Ca, Cout, mask=False):
    D, H, W = shape
    xa = torch.randn((1,) + shape + (Ca,), device=dev).to(torch.bfloat16)
    w = torch.randn((Cout, max(Ca, 8), 3, 3, 3), device=dev) * 0.05
    b = torch.zeros(Cout, device=dev)
    ws_, cps = tc.pack_weights_t(w, variant="s")
    m = torch.randn((1,) + shape + (Cout,), device=dev).to(torch.bfloat16) if mask else None
    res = {}
    bias = None if mask else b
    for cfg in (os.environ.get("ABLATE_CFGS", "tma1_nacc6_epi1,tma1_nacc3_epi1,tma1_nacc6_epi0,tma0_nacc3_epi0")).split(","):
        kv = dict((x[:-1], x[-1]) for x in cfg.split("_"))
        os.environ["VXM_B200_TMA"] = kv.get("tma", "1")
        os.environ["VXM_B200_TCS_NACC"] = kv.get("nacc", "6")
        os.environ["VXM_B200_TCS_EPI"] = kv.get("epi", "1")
        for dbg, label in ((0, "full"), (1, "no_mma"), (4, "no_store"), (7, "loader_only"),
                           (14, "mma_only"), (13, "tmem_ld_only"), (11, "store_only"), (15, "skeleton")):
            os.environ["VXM_B200_TCS_DBG"] = str(dbg)
            t = timeit(lambda: tc.conv_fwd_t(xa, None, ws_, cps, bias, Cout, 3, slope=0.2, mask=m))
            res["%s:%s" % (cfg, label)] = round(t * 1e3, 1)
    for k in ("VXM_B200_TCS_DBG", "VXM_B200_TMA", "VXM_B200_TCS_NACC", "VXM_B200_TCS_EPI"):
        os.environ.pop(k, None)
    print(json.dumps(dict(layer=name, us=res)), flush=True)


which = sys.argv[1:] or ["rem2", "rem1", "rem1d", "flowd"]
if "rem2" in which:
    layer("rem2 16->16", FULL, 16, 16)
if "rem1" in which:
    layer("rem1 32->16", FULL, 32, 16)
if "rem1d" in which:
    layer("rem1 dgrad 16->32 (mask)", FULL, 16, 32, mask=True)
if "flowd" in which:
    layer("rem2 dgrad 16->16 (mask)", FULL, 16, 16, mask=True)
