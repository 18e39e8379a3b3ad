"""Median CUDA-event time of every full-resolution convolution launch of the benchmark step (forward, dgrad, wgrad), under the
environment switches given as KEY=VALUE,... sets on the command line (profiling aid; A/B of kernel variants)."""
import sys, os, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxelmorph_b200 import tc

dev = torch.device("cuda:0")
FULL = (160, 192, 224)
HALF = tuple(s // 2 for s in FULL)


def timeit(fn, n=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda._sleep(200000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(statistics.median(ts) * 1e3, 1)


def rnd(shape, c):
    return torch.randn((1,) + shape + (c,), device=dev).to(torch.bfloat16)


def build():
    L = {}
    w = lambda co, ci: torch.randn((co, ci, 3, 3, 3), device=dev) * 0.05
    # forward layers
    for name, ca, cb, up, co in (("enc0_fwd 8->16", 8, 0, False, 16), ("rem0_fwd 32^+16->32", 32, 16, True, 32), ("rem1_fwd 32->16", 32, 0, False, 16),
                                 ("rem2_fwd 16->16", 16, 0, False, 16)):
        xa = rnd(HALF if up else FULL, ca)
        xb = rnd(FULL, cb) if cb else None
        W = w(co, max(ca + cb, 8) if ca + cb != 8 else 2)
        if ca + cb == 8:
            W = w(co, 2)
        pk, cp = tc.pack_weights_t(W, variant="s")
        b = torch.zeros(co, device=dev)
        L[name] = (lambda xa=xa, xb=xb, pk=pk, cp=cp, b=b, co=co, up=up: tc.conv_fwd_t(xa, xb, pk, cp, b, co, 3, up=up, slope=0.2))
    # flow head: 16 -> 3, fp32 planar out
    x = rnd(FULL, 16); W = w(3, 16); pk, cp = tc.pack_weights_t(W, variant="s"); b3 = torch.zeros(3, device=dev)
    L["flow_fwd 16->3 planar"] = lambda: tc.conv_fwd_t(x, None, pk, cp, b3, 3, 3, out_fp32_planar=True)
    # dgrads (transposed weights): g (Cout ch) -> Cin ch
    for name, cg, cin, mask, split in (("flow_dgrad 8->16 mask", 8, 16, True, None), ("rem2_dgrad 16->16 mask", 16, 16, True, None),
                                       ("rem1_dgrad 16->32 mask", 16, 32, True, None), ("rem0_dgrad 32->48 split", 32, 48, False, 32)):
        g = rnd(FULL, cg)
        W = w(cg if cg != 8 else 3, cin)
        pk, cp = tc.pack_weights_t(W, transposed=True, variant="s")
        m = rnd(FULL, cin) if mask else None
        L[name] = (lambda g=g, pk=pk, cp=cp, cin=cin, m=m, split=split: tc.conv_fwd_t(g, None, pk, cp, None, cin, 3, slope=0.2 if m is not None else None, mask=m, split=split))
    # wgrads
    for name, cx, up, cg in (("flow_wgrad x16 g8", 16, False, 8), ("rem2_wgrad x16 g16", 16, False, 16), ("rem1_wgrad x32 g16", 32, False, 16),
                             ("rem0_wgrad_a x32^ g32", 32, True, 32), ("rem0_wgrad_b x16 g32", 16, False, 32), ("enc0_wgrad x8 g16", 8, False, 16)):
        xx = rnd(HALF if up else FULL, cx)
        g = rnd(FULL, cg)
        cout = 3 if cg == 8 else cg
        cin = 2 if cx == 8 else cx
        L[name] = (lambda xx=xx, g=g, cin=cin, cout=cout, up=up: tc.conv_wgrad(xx, None, g, cin, cout, 3, up=up))
    # kd-folded variants of the two layers with 2 / 3 real channels on one side
    planes2 = [torch.rand((1, 1) + FULL, device=dev) for _ in range(2)]
    planes3 = [torch.randn((1, 1) + FULL, device=dev) for _ in range(3)]
    L["fold: planar_fold_kd x (2 planes -> 8ch)"] = lambda: tc.planar_fold_kd(planes2, 8)
    L["fold: planar_fold_kd g (3 planes -> 16ch)"] = lambda: tc.planar_fold_kd(planes3, 16)
    L["plain: planar_to_ndhwc8 (3 planes)"] = lambda: tc.planar_to_ndhwc8(planes3)
    x3 = tc.planar_fold_kd(planes2, 8)
    W0 = w(16, 2); pk0, cp0 = tc.pack_weights_fold(W0); b16 = torch.zeros(16, device=dev)
    L["fold: enc0_fwd 2D (6 of 8)->16"] = lambda: tc.conv_fwd_t(x3, None, pk0, cp0, b16, 16, 1, slope=0.2)
    g3 = tc.planar_fold_kd(planes3, 16)
    Wf = w(3, 16); pkf, cpf = tc.pack_weights_fold(Wf, transposed=True); m16 = rnd(FULL, 16)
    L["fold: flow_dgrad 2D (9 of 16)->16 mask"] = lambda: tc.conv_fwd_t(g3, None, pkf, cpf, None, 16, 1, slope=0.2, mask=m16)
    batch = tc.WgradBatch.get(dev)
    gw0 = torch.empty((16, 6, 1, 3, 3), device=dev); gb0 = torch.empty(16, device=dev); gz16 = rnd(FULL, 16)
    gwf = torch.empty((9, 16, 1, 3, 3), device=dev); gbf = torch.empty(9, device=dev); x16 = rnd(FULL, 16)

    def khm(x, g, gw, gb, ci, co):
        batch.add_khm(x, g, gw, gb, ci, co)
        batch.flush()
    L["fold: enc0_wgrad khm"] = lambda: khm(x3, gz16, gw0, gb0, 6, 16)
    L["fold: flow_wgrad khm"] = lambda: khm(x16, g3, gwf, gbf, 16, 9)
    return L


if sys.argv[1:2] == ["--once"]:
    # every selected layer once (ncu replays the launch itself): `ncu --set full -k regex:"conv_tcs|wgrad2_kernel" ... tools/conv_layers.py --once`
    for name, fn in build().items():
        if sys.argv[2:] and not any(t in name for t in sys.argv[2:]):
            continue
        fn()
    torch.cuda.synchronize()
    sys.exit(0)
sets = sys.argv[1:] or [""]
L = build()
res = {}
for s in sets:
    kv = dict(x.split("=") for x in s.split(",") if x)
    for k, v in kv.items():
        os.environ[k] = v
    res[s or "default"] = {name: timeit(fn) for name, fn in L.items()}
    for k in kv:
        os.environ.pop(k, None)
names = list(L)
print("%-28s" % "layer" + "".join("%22s" % (s or "default")[-22:] for s in sets))
for n in names:
    print("%-28s" % n + "".join("%22.1f" % res[s or "default"][n] for s in sets))
print("%-28s" % "sum" + "".join("%22.1f" % sum(res[s or "default"].values()) for s in sets))
print(json.dumps(res))
