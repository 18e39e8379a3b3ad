"""GPU parity tests: every kernel, through the C ABI (voxelmorph_b200 modules -> ctypes ->
libvxm_b200.so), against the oracle (oracle/spec_np.py, oracle/ref_torch.py on CPU) and the golden
vectors frozen from the unmodified reference.

Tolerances (stated per north_star): floating point <= 1e-4 relative (max|diff| / max|ref|);
nearest-neighbour warps bit-exact.  Most linear-warp checks are in fact bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases, ref_torch, spec_np

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REL = 1e-4


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def vxm(cuda):
    import voxelmorph_b200 as v
    v._lib.load()
    return v


def g2(x, dev):
    return t(x).to(dev)


# The linear resampler has two coordinate arithmetics (voxelmorph_b200.layers.linear_arith): "exact" replays torch's
# fp32 op sequence (bit-identical to the torch CPU reference), "fast" (default, what bench.py times) computes
# (p + flow) directly.  Every linear-mode test runs under both; FAST_TOL is the stated bound of the fast path
# relative to the reference's value range (north_star tolerance for floating point: 1e-4).
FAST_TOL = 2e-5


@pytest.fixture(params=["fast", "exact"])
def lin(request, monkeypatch):
    monkeypatch.setenv("VXM_B200_LINEAR_ARITH", request.param)
    return request.param


def same(out, ref, lin, what=""):
    if lin == "exact":
        assert np.array_equal(out, ref), "linear path (exact arithmetic) is expected to be bit-exact vs the torch CPU reference " + str(what)
    else:
        e = rel_err(out, ref)
        assert e <= FAST_TOL, (what, e)


# ---------------------------------------------------------------- SpatialTransformer ----------

def test_warp_golden(vxm, cuda, golden, lin):
    g = golden("layers")
    st = vxm.layers.SpatialTransformer((12, 20, 16))
    out = st(g2(g["src"], cuda), g2(g["flow"], cuda)).cpu().numpy()
    assert rel_err(out, g["warp_lin"]) <= REL
    same(out, g["warp_lin"], lin)
    stn = vxm.layers.SpatialTransformer((12, 20, 16), mode="nearest")
    assert np.array_equal(stn(g2(g["lab"], cuda), g2(g["flow"], cuda)).cpu().numpy(), g["warp_near"])
    assert np.array_equal(stn(g2(g["lab"][:1], cuda), g2(g["tie_flow"], cuda)).cpu().numpy(), g["warp_near_tie"])
    same(st(g2(g["src"][:1], cuda), g2(g["tie_flow"], cuda)).cpu().numpy(), g["warp_lin_tie"], lin, "tie")
    # 2-D
    st2 = vxm.layers.SpatialTransformer((20, 28))
    assert rel_err(st2(g2(g["src2"], cuda), g2(g["flow2"], cuda)).cpu().numpy(), g["warp2_lin"]) <= (1e-6 if lin == "exact" else FAST_TOL)
    st2n = vxm.layers.SpatialTransformer((20, 28), mode="nearest")
    assert np.array_equal(st2n(g2(g["lab2"], cuda), g2(g["flow2"], cuda)).cpu().numpy(), g["warp2_near"])


def test_warp_realseg_crop_bit_exact(vxm, cuda, golden):
    g = golden("realseg_crop")
    seg = g["seg"].astype(np.float32)
    stn = vxm.layers.SpatialTransformer(seg.shape[2:], mode="nearest")
    out = stn(g2(seg, cuda), g2(g["flow"], cuda)).cpu().numpy()
    assert np.array_equal(out.astype(np.uint8), g["moved"])
    assert np.array_equal(out, spec_np.warp(seg, g["flow"], mode="nearest"))


@pytest.mark.parametrize("shape,C,sigma", [((9, 11, 13), 1, 3.0), ((16, 8, 40), 4, 20.0), ((33, 31, 65), 2, 1.0)])
def test_warp_vs_oracle_ragged(vxm, cuda, shape, C, sigma, lin):
    src = np.concatenate([cases.smooth_volume(10 + c, shape) for c in range(C)], axis=1)
    flow = cases.smooth_field(77, 3, shape, scale=sigma)   # sigma=20 pushes most samples out of the volume
    st = vxm.layers.SpatialTransformer(shape)
    out = st(g2(src, cuda), g2(flow, cuda)).cpu().numpy()
    ref = spec_np.warp(src, flow)
    same(out, ref, lin, (shape, C, sigma))
    lab = cases.label_volume(5, shape)
    stn = vxm.layers.SpatialTransformer(shape, mode="nearest")
    assert np.array_equal(stn(g2(lab, cuda), g2(flow, cuda)).cpu().numpy(), spec_np.warp(lab, flow, mode="nearest"))


def test_warp_reciprocal_arith_matches_oracle_variant(vxm, cuda, monkeypatch):
    shape = (24, 40, 48)
    lab = cases.label_volume(8, shape)
    flow = cases.smooth_field(9, 3, shape, scale=6.0)
    monkeypatch.setenv("VXM_B200_NEAREST_ARITH", "cuda")
    stn = vxm.layers.SpatialTransformer(shape, mode="nearest")
    out = stn(g2(lab, cuda), g2(flow, cuda)).cpu().numpy()
    assert np.array_equal(out, spec_np.warp(lab, flow, mode="nearest", div="recip"))


def test_warp_full_size_digest_and_properties(vxm, cuda, lin):
    """BASELINE full size 160x192x224: reference digest (nearest, bit-exact) + size-independent properties."""
    d = json.load(open(os.path.join(GOLDEN, "digests.json")))
    full = (160, 192, 224)
    lab = cases.label_volume(101, full)
    flow = cases.smooth_field(102, 3, full, scale=8.0)
    assert hashlib.sha256(lab.tobytes()).hexdigest() == d["lab_sha256"]
    assert hashlib.sha256(flow.tobytes()).hexdigest() == d["flow_sha256"]
    stn = vxm.layers.SpatialTransformer(full, mode="nearest")
    moved = stn(g2(lab, cuda), g2(flow, cuda)).cpu().numpy()
    assert hashlib.sha256(moved.tobytes()).hexdigest() == d["nearest_full_sha256"]
    st = vxm.layers.SpatialTransformer(full)
    vol = cases.smooth_volume(103, full)
    lin = st(g2(vol, cuda), g2(flow, cuda))
    assert abs(float(lin.double().sum()) - d["linear_full_sum"]) <= 1e-6 * d["linear_full_abs_sum"]
    # zero flow: nearest identity exact, linear identity to 1 ulp-ish; linearity in src
    zero = torch.zeros(1, 3, *full, device=cuda)
    assert torch.equal(stn(g2(lab, cuda), zero).cpu(), t(lab))
    v = g2(vol, cuda)
    assert float((st(v, zero) - v).abs().max()) < 2e-5   # coordinate round trip is exact only to ~S*2^-24 voxels
    f = g2(flow, cuda)
    a = st(v, f)
    b = st(2 * v, f)
    assert torch.equal(b, 2 * a)   # scaling by 2 commutes with every rounding
    # integer shift along W with zero fill
    sh = torch.zeros(1, 3, *full, device=cuda)
    sh[:, 2] = 3.0
    o = stn(g2(lab, cuda), sh).cpu()
    assert torch.equal(o[..., :-3], t(lab)[..., 3:]) and not o[..., -3:].any()


def test_warp_backward_vs_autograd(vxm, cuda, lin):
    shape = (10, 12, 14)
    src = np.concatenate([cases.smooth_volume(1, shape), cases.smooth_volume(2, shape)], axis=1)
    flow = cases.smooth_field(3, 3, shape, scale=3.0)
    gout = cases.smooth_field(4, 2, shape, scale=1.0)
    s_c = t(src).double().requires_grad_(True)
    f_c = t(flow).double().requires_grad_(True)
    ref_torch.spatial_transform(s_c, f_c).backward(t(gout).double())
    s_g = g2(src, cuda).requires_grad_(True)
    f_g = g2(flow, cuda).requires_grad_(True)
    vxm.layers.SpatialTransformer(shape)(s_g, f_g).backward(g2(gout, cuda))
    assert rel_err(s_g.grad.cpu().numpy(), s_c.grad.numpy()) <= REL
    assert rel_err(f_g.grad.cpu().numpy(), f_c.grad.numpy()) <= REL
    # 2-D
    s2, f2 = cases.smooth_volume(5, (18, 22)), cases.smooth_field(6, 2, (18, 22), scale=2.0)
    s_c = t(s2).double().requires_grad_(True)
    f_c = t(f2).double().requires_grad_(True)
    ref_torch.spatial_transform(s_c, f_c).sum().backward()
    s_g = g2(s2, cuda).requires_grad_(True)
    f_g = g2(f2, cuda).requires_grad_(True)
    vxm.layers.SpatialTransformer((18, 22))(s_g, f_g).sum().backward()
    assert rel_err(s_g.grad.cpu().numpy(), s_c.grad.numpy()) <= REL
    assert rel_err(f_g.grad.cpu().numpy(), f_c.grad.numpy()) <= REL


# ---------------------------------------------------------------- VecInt -----------------------

def test_vecint_golden(vxm, cuda, golden, lin):
    g = golden("layers")
    for n in (0, 1, 4, 7):
        out = vxm.layers.VecInt((12, 20, 16), n)(g2(g["vel"], cuda)).cpu().numpy()
        same(out, g["vecint_%d" % n], lin, n)
    out = vxm.layers.VecInt((20, 28), 5)(g2(g["vel2"], cuda)).cpu().numpy()
    assert rel_err(out, g["vecint2_5"]) <= 1e-6


def test_vecint_train_path_and_backward(vxm, cuda, lin):
    shape = (10, 12, 14)
    vel = cases.smooth_field(31, 3, shape, scale=4.0)
    gout = cases.smooth_field(32, 3, shape, scale=1.0)
    for n in (0, 1, 3, 7):
        v_c = t(vel).double().requires_grad_(True)
        out_c = ref_torch.vec_int(v_c, n)
        out_c.backward(t(gout).double())
        v_g = g2(vel, cuda).requires_grad_(True)
        out_g = vxm.layers.VecInt(shape, n)(v_g)
        out_g.backward(g2(gout, cuda))
        same(out_g.detach().cpu().numpy(), spec_np.vecint(vel, n), lin, n)   # states path = ping-pong path
        with torch.no_grad():
            assert torch.equal(vxm.layers.VecInt(shape, n)(g2(vel, cuda)), out_g.detach()), n
        assert rel_err(v_g.grad.cpu().numpy(), v_c.grad.numpy()) <= REL, n
    v2 = cases.smooth_field(33, 2, (18, 22), scale=3.0)
    v_c = t(v2).double().requires_grad_(True)
    ref_torch.vec_int(v_c, 4).sum().backward()
    v_g = g2(v2, cuda).requires_grad_(True)
    vxm.layers.VecInt((18, 22), 4)(v_g).sum().backward()
    assert rel_err(v_g.grad.cpu().numpy(), v_c.grad.numpy()) <= REL


def test_vecint_full_size_properties(vxm, cuda, lin):
    """Half-res benchmark size 80x96x112 and the 128^3 sweep size: oracle equality + VecInt(0) identity."""
    shape = (80, 96, 112)
    vel = cases.smooth_field(41, 3, shape, scale=5.0)
    out = vxm.layers.VecInt(shape, 7)(g2(vel, cuda)).cpu().numpy()
    same(out, spec_np.vecint(vel, 7), lin)
    v = g2(vel, cuda)
    assert torch.equal(vxm.layers.VecInt(shape, 0)(v), v)
    big = torch.randn(2, 3, 128, 128, 128, device=cuda)
    o1 = vxm.layers.VecInt((128,) * 3, 5)(big)
    o2 = vxm.layers.VecInt((128,) * 3, 5)(big[1:])
    assert torch.equal(o1[1:], o2)   # batch entries are independent


# ---------------------------------------------------------------- ResizeTransform --------------

def test_resize_golden_and_backward(vxm, cuda, golden):
    g = golden("layers")
    for key, x, vr, nd in (("resize_down", "flow", 2, 3), ("resize_up", "flow", 0.5, 3), ("resize_one", "flow", 1, 3),
                           ("resize_down_odd", "odd", 2, 3), ("resize_up_odd", "odd", 0.5, 3),
                           ("resize2_down", "flow2", 2, 2), ("resize2_up", "flow2", 0.5, 2)):
        out = vxm.layers.ResizeTransform(vr, nd)(g2(g[x], cuda)).cpu().numpy()
        assert out.shape == g[key].shape, key
        assert rel_err(out, g[key]) <= 2e-6, key
    for x, vr, nd in ((g["flow"], 2, 3), (g["flow"], 0.5, 3), (g["odd"], 2, 3), (g["odd"], 0.5, 3), (g["flow2"], 2, 2),
                      (g["flow2"], 0.5, 2), (g["odd"], 3, 3)):
        x_c = t(x).double().requires_grad_(True)
        o_c = ref_torch.resize_transform(x_c, vr)
        w = torch.from_numpy(np.random.default_rng(0).standard_normal(tuple(o_c.shape)))
        (o_c * w).sum().backward()
        x_g = g2(x, cuda).requires_grad_(True)
        o_g = vxm.layers.ResizeTransform(vr, nd)(x_g)
        assert tuple(o_g.shape) == tuple(o_c.shape)
        assert rel_err(o_g.detach().cpu().numpy(), o_c.detach().numpy()) <= REL
        (o_g * w.float().to(cuda)).sum().backward()
        assert rel_err(x_g.grad.cpu().numpy(), x_c.grad.numpy()) <= REL, (vr, nd)


def test_resize_full_size_roundtrip_property(vxm, cuda):
    full = (160, 192, 224)
    f = g2(cases.smooth_field(51, 3, full, scale=4.0), cuda)
    down = vxm.layers.ResizeTransform(2, 3)(f)
    assert tuple(down.shape) == (1, 3, 80, 96, 112)
    up = vxm.layers.ResizeTransform(0.5, 3)(down)
    assert tuple(up.shape) == (1, 3) + full
    # a field that is linear in the coordinates is reproduced exactly by linear resampling (up to scale)
    lin = torch.zeros(1, 3, *full, device=cuda)
    lin[:, 0] = torch.arange(160, device=cuda).float()[:, None, None] * 0.5
    d2 = vxm.layers.ResizeTransform(2, 3)(lin)
    exp = torch.arange(80, device=cuda).float() * (159.0 / 79.0) * 0.25
    assert float((d2[0, 0, :, 0, 0] - exp).abs().max()) < 1e-4


# ---------------------------------------------------------------- losses -----------------------

def test_losses_golden(vxm, cuda, golden):
    g = golden("losses")
    L = vxm.losses

    def run(fn, pred):
        p = g2(pred, cuda).requires_grad_(True)
        v = fn(p)
        v.backward()
        return float(v.item()), p.grad.cpu().numpy()

    v, gr = run(lambda p: L.NCC().loss(g2(g["I"], cuda), p), g["J"])
    assert abs(v - g["ncc"]) <= REL * abs(g["ncc"])
    assert rel_err(gr, spec_np.ncc_grad_pred(g["I"], g["J"])) <= 1e-3     # fp32 cancellation in the variance terms
    assert rel_err(gr, g["ncc_grad"]) <= 2e-3                              # (the fp32 reference itself is ~1e-3 from fp64)
    v, gr = run(lambda p: L.NCC(win=[5, 5, 5]).loss(g2(g["I"], cuda), p), g["J"])
    assert abs(v - g["ncc5"]) <= REL * abs(g["ncc5"])
    assert rel_err(gr, spec_np.ncc_grad_pred(g["I"], g["J"], [5, 5, 5])) <= 1e-3
    v, gr = run(lambda p: L.NCC().loss(g2(g["I2"], cuda), p), g["J2"])
    assert abs(v - g["ncc2"]) <= REL * abs(g["ncc2"])
    assert rel_err(gr, spec_np.ncc_grad_pred(g["I2"], g["J2"])) <= 1e-3
    v, gr = run(lambda p: L.MSE().loss(g2(g["I"], cuda), p), g["J"])
    assert abs(v - g["mse"]) <= 1e-6 * abs(g["mse"]) and rel_err(gr, g["mse_grad"]) <= 1e-6
    v, gr = run(lambda p: L.Grad("l2", loss_mult=2).loss(None, p), g["gflow"])
    assert abs(v - g["grad_l2"]) <= 1e-6 * abs(g["grad_l2"]) and rel_err(gr, g["grad_l2_grad"]) <= 1e-5
    v, gr = run(lambda p: L.Grad("l1").loss(None, p), g["gflow"])
    assert abs(v - g["grad_l1"]) <= 1e-6 * abs(g["grad_l1"]) and rel_err(gr, g["grad_l1_grad"]) <= 1e-5
    v, gr = run(lambda p: L.Dice().loss(g2(g["dice_true"], cuda), p), g["dice_pred"])
    assert abs(v - g["dice"]) <= 1e-6 and rel_err(gr, g["dice_grad"]) <= 1e-5


def test_ncc_full_size_properties(vxm, cuda):
    full = (160, 192, 224)
    I, J = cases.volume_pair(61, full, sigma=3.0)
    Ig, Jg = g2(I, cuda), g2(J, cuda)
    ncc = vxm.losses.NCC()
    a = float(ncc.loss(Ig, Jg).item())
    b = float(ncc.loss(Jg, Ig).item())
    assert abs(a - b) <= 1e-5 * abs(a)               # cc is symmetric in (I, J)
    assert -1.0 - 1e-3 <= a <= 0.0
    same = float(ncc.loss(Ig, Ig).item())
    assert same < a and same < -0.9                   # NCC(I, I) -> -(fraction of windows with variance)
    z = torch.zeros_like(Ig)
    assert float(ncc.loss(z, z).item()) == 0.0
    ref = spec_np.ncc_loss(I[..., :48, :64, :64], J[..., :48, :64, :64])
    sub = float(ncc.loss(Ig[..., :48, :64, :64].contiguous(), Jg[..., :48, :64, :64].contiguous()).item())
    assert abs(sub - ref) <= REL * abs(ref)


def test_loss_errors(vxm, cuda):
    with pytest.raises(AssertionError):
        vxm.losses.Grad("l3").loss(None, torch.zeros(1, 3, 4, 4, 4, device=cuda))
    with pytest.raises(vxm._lib.VxmError):
        vxm.losses.MSE().loss(torch.zeros(4), torch.zeros(4))     # CPU tensors: loud failure, no fallback


# ---------------------------------------------------------------- U-Net pieces -----------------

@pytest.mark.parametrize("shape,cin,cout,slope", [((6, 7, 9), 3, 5, 0.2), ((16, 16, 32), 16, 32, 0.2), ((8, 24, 40), 48, 32, 0.2),
                                                   ((8, 8, 8), 16, 3, None), ((20, 36), 2, 16, 0.2), ((12, 12), 8, 2, None)])
def test_conv_fwd_bwd(vxm, cuda, shape, cin, cout, slope):
    from voxelmorph_b200 import ops
    nd = len(shape)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn((2, cin) + shape, generator=gen)
    w = torch.randn((cout, cin) + (3,) * nd, generator=gen) * 0.1
    b = torch.randn(cout, generator=gen)
    gy = torch.randn((2, cout) + shape, generator=gen)
    conv = torch.nn.functional.conv3d if nd == 3 else torch.nn.functional.conv2d
    xc, wc, bc = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yc = conv(xc, wc, bc, padding=1)
    if slope is not None:
        yc = torch.nn.functional.leaky_relu(yc, slope)
    yc.backward(gy.double())
    xg, wg, bg = x.to(cuda).requires_grad_(True), w.to(cuda).requires_grad_(True), b.to(cuda).requires_grad_(True)
    yg = ops.conv_k3(xg, wg, bg, slope)
    yg.backward(gy.to(cuda))
    assert rel_err(yg.detach().cpu().numpy(), yc.detach().numpy()) <= 1e-5
    assert rel_err(xg.grad.cpu().numpy(), xc.grad.numpy()) <= 1e-5
    assert rel_err(wg.grad.cpu().numpy(), wc.grad.numpy()) <= 1e-5
    assert rel_err(bg.grad.cpu().numpy(), bc.grad.numpy()) <= 1e-5
    # small case also against the explicit numpy restatement
    if np.prod(shape) < 1000:
        ref = spec_np.conv_k3(x.numpy(), w.numpy(), b.numpy(), slope)
        assert rel_err(yg.detach().cpu().numpy(), ref) <= 1e-5


@pytest.mark.parametrize("shape", [(4, 6, 8), (16, 32, 16), (10, 12)])
def test_pool_upcat(vxm, cuda, shape):
    from voxelmorph_b200 import ops
    nd = len(shape)
    gen = torch.Generator().manual_seed(2)
    x = torch.randn((2, 3) + shape, generator=gen)
    pool = torch.nn.functional.max_pool3d if nd == 3 else torch.nn.functional.max_pool2d
    xc = x.clone().requires_grad_(True)
    yc = pool(xc, 2)
    gy = torch.randn(yc.shape, generator=gen)
    yc.backward(gy)
    xg = x.to(cuda).requires_grad_(True)
    yg = ops.maxpool2(xg)
    yg.backward(gy.to(cuda))
    assert torch.equal(yg.detach().cpu(), yc.detach())
    assert torch.equal(xg.grad.cpu(), xc.grad)
    assert np.array_equal(yg.detach().cpu().numpy(), spec_np.maxpool2(x.numpy()))
    skip = torch.randn((2, 5) + tuple(2 * s for s in shape), generator=gen)
    ac, sc = x.clone().requires_grad_(True), skip.clone().requires_grad_(True)
    oc = torch.cat([torch.nn.functional.interpolate(ac, scale_factor=2, mode="nearest"), sc], dim=1)
    go = torch.randn(oc.shape, generator=gen)
    oc.backward(go)
    ag, sg = x.to(cuda).requires_grad_(True), skip.to(cuda).requires_grad_(True)
    og = ops.upsample2_cat(ag, sg)
    og.backward(go.to(cuda))
    assert torch.equal(og.detach().cpu(), oc.detach())
    assert rel_err(ag.grad.cpu().numpy(), ac.grad.numpy()) <= 1e-6
    assert torch.equal(sg.grad.cpu(), sc.grad)


def test_adam_parity(vxm, cuda):
    gen = torch.Generator().manual_seed(3)
    ps = [torch.randn(5, 7, generator=gen), torch.randn(11, generator=gen)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    mine = [torch.nn.Parameter(p.clone().to(cuda)) for p in ps]
    ropt = torch.optim.Adam(ref, lr=1e-2)
    mopt = vxm.optim.FusedAdam(mine, lr=1e-2)
    for step in range(5):
        gs = [torch.randn(p.shape, generator=gen) for p in ps]
        for p, gr in zip(ref, gs):
            p.grad = gr.clone()
        mopt.zero_grad()
        for p, gr in zip(mine, gs):
            p.grad.copy_(gr.to(cuda))
        ropt.step()
        mopt.step()
        for p, q in zip(ref, mine):
            assert rel_err(q.detach().cpu().numpy(), p.detach().numpy()) <= 1e-6, step


def test_resize_upsampling_adjoint_kernels_agree(vxm, cuda, monkeypatch):
    """The shared-memory column-marching adjoint (default for x2 upsampling) against the plain marching kernel and fp64 autograd,
    on a ragged size whose tiles are partial on every axis."""
    from voxelmorph_b200 import _lib
    lib = _lib.load()
    Di, Hi, Wi = 21, 27, 45
    Do, Ho, Wo = 42, 54, 90
    g = torch.Generator().manual_seed(5)
    go = torch.randn((1, 3, Do, Ho, Wo), generator=g).to(cuda)
    outs = {}
    for mode in ("", "march"):
        if mode:
            monkeypatch.setenv("VXM_B200_RESIZE_BWD", mode)
        gx = torch.empty((1, 3, Di, Hi, Wi), device=cuda)
        assert lib.vxm_resize_bwd(_lib.ptr(go), _lib.ptr(gx), 1, 3, Di, Hi, Wi, Do, Ho, Wo, 2.0, 1.0, _lib.stream_ptr()) == 0, _lib.last_error()
        outs[mode] = gx.cpu()
    x = torch.zeros((1, 3, Di, Hi, Wi), dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.interpolate(x * 2.0, size=(Do, Ho, Wo), mode="trilinear", align_corners=True)
    (y * go.cpu().double()).sum().backward()
    errs = {mode: rel_err(v.numpy(), x.grad.numpy()) for mode, v in outs.items()}
    errs["mutual"] = rel_err(outs[""].numpy(), outs["march"].numpy())
    print("resize adjoint errors:", errs)
    assert max(errs.values()) <= 1e-5, errs
