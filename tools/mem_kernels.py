"""Launch each memory-bound kernel a few times at the benchmark shapes (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import voxelmorph_b200 as vxm
dev = torch.device("cuda:0")
shape = (160, 192, 224)
half = tuple(s // 2 for s in shape)
src = torch.rand((1, 1) + shape, device=dev)
flow = (torch.randn((1, 3) + shape, device=dev) * 3.0).requires_grad_(True)
st = vxm.layers.SpatialTransformer(shape)
vel = (torch.randn((1, 3) + half, device=dev) * 2.0).requires_grad_(True)
vi = vxm.layers.VecInt(half, 7)
down, up = vxm.layers.ResizeTransform(2, 3), vxm.layers.ResizeTransform(0.5, 3)
I, J = torch.rand((1, 1) + shape, device=dev), torch.rand((1, 1) + shape, device=dev).requires_grad_(True)
ncc = vxm.losses.NCC().loss
for _ in range(3):
    y = st(src, flow); y.sum().backward()
    o = vi(vel); o.sum().backward()
    d = down(flow); d.sum().backward()
    u = up(vel); u.sum().backward()
    l = ncc(I, J); l.backward()
torch.cuda.synchronize()
