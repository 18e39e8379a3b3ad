"""CPU check of the index algebra behind the kd-folded layers (voxelmorph_b200/engine_bf16.py): folding the
three kd taps into the channels turns the reference's 3-D convolutions (voxelmorph/torch/networks.py:299 first ConvBlock,
:211 flow head) into 2-D ones per slice; forward, data gradient and weight / bias gradients must be the same numbers.  Everything
here is torch on the CPU in float64: it pins the maps the CUDA kernels implement (tests/test_gpu_tc.py checks the kernels)."""
import torch
import torch.nn.functional as F

from voxelmorph_b200 import engine_bf16 as E


def conv2d_per_slice(x_cl, w2):
    """x_cl (B, D, H, W, C) channels-last, w2 (Cout, C, 3, 3): 2-D cross-correlation of every slice, zero padding 1."""
    B, D, H, W, C = x_cl.shape
    y = F.conv2d(x_cl.permute(0, 1, 4, 2, 3).reshape(B * D, C, H, W), w2, padding=1)
    return y.reshape(B, D, -1, H, W).permute(0, 2, 1, 3, 4)            # (B, Cout, D, H, W)


def test_first_layer_forward_and_weight_gradient():
    g = torch.Generator().manual_seed(1)
    B, P, Cout, shape = 2, 2, 5, (6, 7, 9)
    x = torch.randn((B, P) + shape, generator=g, dtype=torch.float64)
    w = torch.randn((Cout, P, 3, 3, 3), generator=g, dtype=torch.float64, requires_grad=True)
    ref = F.conv3d(x, w, padding=1)
    x3 = E.fold_planes([x[:, i:i + 1] for i in range(P)])
    assert x3.shape == (B,) + shape + (3 * P,)
    w2 = E.fold_weight_first(w.detach()).requires_grad_(True)
    out = conv2d_per_slice(x3, w2)
    assert torch.allclose(out, ref, atol=1e-12)
    gy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    ref.backward(gy)
    out.backward(gy)
    gw = E.unfold_grad_first(w2.grad.reshape(Cout, 3 * P, 1, 3, 3), Cout, P)
    assert torch.allclose(gw, w.grad, atol=1e-10)


def test_flow_head_data_and_weight_gradients():
    g = torch.Generator().manual_seed(2)
    B, Cin, nd, shape = 1, 4, 3, (5, 6, 8)
    x = torch.randn((B, Cin) + shape, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn((nd, Cin, 3, 3, 3), generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(nd, dtype=torch.float64, requires_grad=True)
    gy = torch.randn((B, nd) + shape, generator=g, dtype=torch.float64)
    F.conv3d(x, w, b, padding=1).backward(gy)
    g3 = E.fold_planes([gy[:, i:i + 1] for i in range(nd)])             # (B, D, H, W, 3 nd): channel kd' * nd + c
    # data gradient = 2-D convolution of the folded flow gradient with the flipped, folded weights
    dx = conv2d_per_slice(g3, E.fold_weight_flow_dgrad(w.detach()))
    assert torch.allclose(dx, x.grad, atol=1e-10)
    # weight gradient: the 2-D gradient of "x (2-D conv) -> 3 nd channels" with g3 as the output gradient, mapped back
    B_, D, H, W, _ = g3.shape
    x2 = x.detach().permute(0, 2, 1, 3, 4).reshape(B_ * D, Cin, H, W)
    wd = torch.zeros((3 * nd, Cin, 3, 3), dtype=torch.float64, requires_grad=True)
    bd = torch.zeros(3 * nd, dtype=torch.float64, requires_grad=True)
    F.conv2d(x2, wd, bd, padding=1).backward(g3.permute(0, 1, 4, 2, 3).reshape(B_ * D, 3 * nd, H, W))
    gw, gb = E.unfold_grad_flow(wd.grad.reshape(3 * nd, Cin, 1, 3, 3), bd.grad, nd, Cin)
    assert torch.allclose(gw, w.grad, atol=1e-10)
    assert torch.allclose(gb, b.grad, atol=1e-10)
