"""A/B of the fast VecInt backward's debug variants (VXM_B200_VECINT_DBG bits: 1 no warp pre-reduction, 2 interleaved blocks,
4 one voxel per thread).  Profiling aid."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxelmorph_b200 import _lib
lib = _lib.load(); P, S = _lib.ptr, _lib.stream_ptr
dev = torch.device("cuda:0")
Dh, Hh, Wh = 80, 96, 112
vel = torch.nn.functional.interpolate(torch.randn((1, 3, 5, 6, 7), device=dev) * 2.0, size=(Dh, Hh, Wh), mode="trilinear", align_corners=True).contiguous()
g3h = torch.rand_like(vel); out = torch.empty_like(vel)
states = torch.empty(int(lib.vxm_vecint_fast_states_bytes(1, Dh, Hh, Wh, 7)), dtype=torch.uint8, device=dev)
work_b = torch.empty(int(lib.vxm_vecint_fast_work_bytes(1, Dh, Hh, Wh, 1)), dtype=torch.uint8, device=dev)
assert lib.vxm_vecint_fwd(P(vel), P(out), P(states), None, 1, Dh, Hh, Wh, 3, 7, 2, S()) == 0
ref = None
for dbg in range(8):
    os.environ["VXM_B200_VECINT_DBG"] = str(dbg)
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000); a.record()
        assert lib.vxm_vecint_bwd(P(g3h), P(states), P(out), P(work_b), 1, Dh, Hh, Wh, 3, 7, 2, S()) == 0
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    if ref is None:
        ref = out.clone()
    print("dbg %d: %7.1f us   max dev vs dbg0 %.2e" % (dbg, statistics.median(ts[1:]), float((out - ref).abs().max())), flush=True)
