"""GPU tests of the tensor-core (bf16) U-Net engine: Unet + flow head forward and the hand-written backward against the
CPU oracle with bf16 storage emulated at the same points (oracle/ref_torch.emulate_bf16), and the measured deviation from
the pure-fp32 oracle (reported; bf16 operands cannot meet the 1e-4 parity the fp32 engine meets)."""
import numpy as np
import pytest
import torch

from oracle import cases, ref_torch

from test_oracle import full_cfg

pytestmark = pytest.mark.gpu

F16 = [[16, 16, 16, 16], [16, 16, 16, 16, 16, 16, 16]]
BF16_VARIANTS = {
    "default3d": dict(inshape=(32, 32, 48)),
    "feat16_3d": dict(inshape=(16, 32, 32), nb_unet_features=F16),
    "halfres3d": dict(inshape=(16, 16, 32), unet_half_res=True),
    "ncpl2_3d": dict(inshape=(16, 16, 16), nb_unet_features=16, nb_unet_levels=3, unet_feat_mult=1, nb_unet_conv_per_level=2),
    "config1_2d": dict(inshape=(64, 64), int_steps=0),
    "bidir2d": dict(inshape=(32, 48), bidir=True, int_steps=5),
}


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture()
def vxm_bf16(cuda, monkeypatch):
    import voxelmorph_b200 as v
    v._lib.load()
    monkeypatch.setenv("VXM_B200_CONV_ENGINE", "bf16")
    yield v
    ref_torch.emulate_bf16(False)


@pytest.mark.parametrize("name", sorted(BF16_VARIANTS))
def test_bf16_engine_forward_backward(vxm_bf16, cuda, name):
    vxm = vxm_bf16
    kw = BF16_VARIANTS[name]
    cfg = full_cfg(kw)
    model = vxm.networks.VxmDense(**kw)
    sd = ref_torch.init_state_dict(cfg, seed=77, flow_std=2e-2)
    model.load_state_dict(sd, strict=False)
    model.to(cuda).train()
    shape = kw["inshape"]
    s, tr = cases.volume_pair(93, shape, sigma=1.5)
    S, T = t(s).to(cuda), t(tr).to(cuda)
    out = model(S, T)
    flow = out[-1]
    gen = torch.Generator().manual_seed(1)
    gflow = torch.randn(flow.shape, generator=gen)
    gy = torch.randn(out[0].shape, generator=gen)
    loss = (flow * gflow.to(cuda)).sum() + (out[0] * gy.to(cuda)).sum()
    loss.backward()
    # oracle with bf16 storage emulation
    ref_torch.emulate_bf16(True)
    sdc = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outc = ref_torch.vxm_forward(sdc, cfg, t(s), t(tr))
    ((outc[-1] * gflow).sum() + (outc[0] * gy).sum()).backward()
    ref_torch.emulate_bf16(False)
    with torch.no_grad():
        out32 = ref_torch.vxm_forward(sd, cfg, t(s), t(tr))
    e_flow, e_moved = rel(flow.detach().cpu(), outc[-1].detach()), rel(out[0].detach().cpu(), outc[0].detach())
    d_flow, d_moved = rel(flow.detach().cpu(), out32[-1]), rel(out[0].detach().cpu(), out32[0])
    print("\n[%s] vs bf16-emulating oracle: flow %.2e moved %.2e | vs fp32 oracle: flow %.2e moved %.2e"
          % (name, e_flow, e_moved, d_flow, d_moved))
    assert e_flow <= 2e-2 and e_moved <= 2e-2
    assert d_flow <= 6e-2 and d_moved <= 6e-2
    errs = sorted(((rel(p.grad.cpu(), sdc[k].grad), k) for k, p in model.named_parameters()), reverse=True)
    print("[%s] worst parameter-gradient rel errs vs emulating oracle: %s" % (name, ", ".join("%s %.2e" % (k, e) for e, k in errs[:3])))
    # bf16 gradient storage costs ~0.4% per layer; the tiny deepest levels (a handful of voxels) are the noisiest
    assert errs[0][0] <= 1.5e-1, (name, errs[:3])
    assert np.median([e for e, _ in errs]) <= 5e-2


def test_bf16_engine_train_step_tracks_fp32(vxm_bf16, cuda, golden):
    """One full training step (NCC + Grad, Adam) in bf16 mode lands within 1e-2 of the fp32 reference loss."""
    vxm = vxm_bf16
    g = golden("vxmdense")
    kw = dict(inshape=(32, 32, 48))
    cfg = full_cfg(kw)
    model = vxm.networks.VxmDense(**kw)
    model.load_state_dict(ref_torch.init_state_dict(cfg, seed=1234, flow_std=2e-2), strict=False)
    model.to(cuda).train()
    s, tr = cases.volume_pair(91, kw["inshape"], sigma=1.5)
    S, T = t(s).to(cuda), t(tr).to(cuda)
    opt = vxm.optim.FusedAdam(model.parameters(), lr=1e-4)
    opt.zero_grad()
    y, flow = model(S, T)
    loss = vxm.losses.NCC().loss(T, y) + 0.01 * vxm.losses.Grad("l2", loss_mult=2).loss(None, flow)
    loss.backward()
    opt.step()
    ref = float(g["default3d/loss"])
    assert abs(float(loss.item()) - ref) <= 1e-2 * abs(ref)
    # second step exercises the refreshed packed weights
    opt.zero_grad()
    y, flow = model(S, T)
    loss2 = vxm.losses.NCC().loss(T, y) + 0.01 * vxm.losses.Grad("l2", loss_mult=2).loss(None, flow)
    loss2.backward()
    opt.step()
    assert np.isfinite(float(loss2.item()))


def test_bf16_engine_rejects_unsupported_shapes(vxm_bf16, cuda):
    vxm = vxm_bf16
    m = vxm.networks.VxmDense((16, 16, 16), nb_unet_features=[[4, 8, 8, 8], [8, 8, 8, 8, 8, 4, 4]]).to(cuda)
    with pytest.raises(vxm._lib.VxmError, match="VXM_B200_CONV_ENGINE=f32"):
        m(torch.rand(1, 1, 16, 16, 16, device=cuda), torch.rand(1, 1, 16, 16, 16, device=cuda))


def test_graphed_train_step_matches_eager(vxm_bf16, cuda):
    """The CUDA-graph replay of the whole step produces the same loss sequence / parameters as eager launches."""
    vxm = vxm_bf16
    from voxelmorph_b200.trainer import GraphedTrainStep
    kw = dict(inshape=(32, 32, 32))
    cfg = full_cfg(kw)
    s, tr = cases.volume_pair(95, kw["inshape"], sigma=1.5)
    S, T = t(s).to(cuda), t(tr).to(cuda)

    def make():
        m = vxm.networks.VxmDense(**kw)
        m.load_state_dict(ref_torch.init_state_dict(cfg, seed=5, flow_std=2e-2), strict=False)
        m.to(cuda).train()
        return m, vxm.optim.FusedAdam(m.parameters(), lr=1e-3)

    m1, o1 = make()
    eager = []
    for _ in range(6):
        o1.zero_grad()
        y, flow = m1(S, T)
        loss = vxm.losses.NCC().loss(T, y) + 0.01 * vxm.losses.Grad("l2", loss_mult=2).loss(None, flow)
        loss.backward()
        o1.step()
        eager.append(float(loss))
    m2, o2 = make()
    tr2 = GraphedTrainStep(m2, o2, warmup=3).capture(S, T)      # 3 eager warm-up steps; capture itself executes nothing
    graphed = [float(tr2(S, T)) for _ in range(3)]
    # replays are steps 4, 5, 6 of the trajectory (atomics in the VecInt / warp backward make the two runs agree only
    # to rounding)
    for i in range(3):
        assert abs(graphed[i] - eager[3 + i]) <= 2e-3 * abs(eager[3 + i]), (i, graphed, eager)
    assert int(o2.step_dev.item()) == 6
