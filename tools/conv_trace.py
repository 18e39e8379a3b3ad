"""In-kernel clock64() trace of the kw-stacked convolution's pipeline roles (profiling aid): prints, for CTA 0, the median
cycle cost of every stage of the MMA issuers, the slab producer and the epilogue groups."""
import sys, os, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxelmorph_b200 import tc

dev = torch.device("cuda:0")
FULL = (160, 192, 224)
Ca, Cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 16)
xa = torch.randn((1,) + FULL + (Ca,), device=dev).to(torch.bfloat16)
w = torch.randn((Cout, Ca, 3, 3, 3), device=dev) * 0.05
b = torch.zeros(Cout, device=dev)
ws_, cps = tc.pack_weights_t(w, variant="s")
for dbg in [int(x) for x in os.environ.get("TRACE_DBG", "0,15,14").split(",")]:
    trace = torch.zeros(6 * 4096, dtype=torch.int64, device=dev)
    os.environ["VXM_B200_TCS_DBG"] = str(dbg)
    tc.conv_fwd_t(xa, None, ws_, cps, b, Cout, 3, slope=0.2)       # warm
    os.environ["VXM_B200_TCS_TRACE"] = str(trace.data_ptr())
    tc.conv_fwd_t(xa, None, ws_, cps, b, Cout, 3, slope=0.2)
    torch.cuda.synchronize()
    os.environ.pop("VXM_B200_TCS_TRACE")
    t = trace.cpu().view(6, 512, 8).numpy()
    out = {"dbg": dbg}
    names = {0: ["start", "slabs_seen", "acc0_free", "half0_issued", "acc1_free", "half1_issued"],
             2: ["start", "slot_free", "issued"], 3: ["start", "tfull", "tmem_read+released", "stored"]}
    for role, nm in ((0, "issuer0"), (1, "issuer1"), (2, "producer"), (3, "epi0"), (4, "epi1"), (5, "epi2")):
        pts = names[0 if role < 2 else (2 if role == 2 else 3)]
        rows = [r for r in t[role][8:200] if r[0] > 0]
        if len(rows) < 4:
            continue
        d = {}
        for k in range(1, len(pts)):
            d[pts[k]] = int(statistics.median([int(r[k]) - int(r[k - 1]) for r in rows if r[k] > 0 and r[k - 1] > 0]))
        starts = [int(r[0]) for r in rows]
        d["period"] = int(statistics.median([starts[i + 1] - starts[i] for i in range(len(starts) - 1)]))
        out[nm] = d
    print(json.dumps(out), flush=True)
os.environ.pop("VXM_B200_TCS_DBG", None)
