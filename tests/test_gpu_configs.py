"""BASELINE.json configs 4 and 5 as parity tests (they are test cases, not bench lines):
  config 4 — 3-D 256^3 inference-only SpatialTransformer + VecInt (sizes the numpy oracle still finishes in seconds);
  config 5 — semi-supervised composition: flow rescaled to the segmentation resolution, LINEAR warp of
             the one-hot segmentation, Dice loss (+ autograd), and the nearest-neighbour label-map warp, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import cases, ref_torch, spec_np

pytestmark = pytest.mark.gpu


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def vxm(cuda):
    import voxelmorph_b200 as v
    v._lib.load()
    return v


@pytest.mark.parametrize("arith", ["fast", "exact"])
def test_config4_256cubed_warp_and_vecint(vxm, cuda, arith, monkeypatch):
    monkeypatch.setenv("VXM_B200_LINEAR_ARITH", arith)   # exact: bit-identical to the oracle; fast (default, benched): <= 2e-5
    full = (256, 256, 256)
    vol = cases.smooth_volume(301, full)
    lab = cases.label_volume(302, full)
    flow = cases.smooth_field(303, 3, full, scale=6.0)
    F = t(flow).to(cuda)
    lin = vxm.layers.SpatialTransformer(full)(t(vol).to(cuda), F).cpu().numpy()
    near = vxm.layers.SpatialTransformer(full, mode="nearest")(t(lab).to(cuda), F).cpu().numpy()
    # oracle on a 64-slice slab of the output (the gather may reach anywhere in the source volume)
    sl = slice(96, 160)
    ref_lin = spec_np.warp(vol, flow)[:, :, sl]
    if arith == "exact":
        assert np.array_equal(lin[:, :, sl], ref_lin)
    else:
        assert rel(lin[:, :, sl], ref_lin) <= 2e-5
    assert np.array_equal(near[:, :, sl], spec_np.warp(lab, flow, mode="nearest")[:, :, sl])
    # VecInt at 128^3 (int_downsize = 2 of a 256^3 volume), all steps in one launch, against the oracle
    half = (128, 128, 128)
    vel = cases.smooth_field(304, 3, half, scale=4.0)
    out = vxm.layers.VecInt(half, 7)(t(vel).to(cuda)).cpu().numpy()
    if arith == "exact":
        assert np.array_equal(out, spec_np.vecint(vel, 7))
    else:
        assert rel(out, spec_np.vecint(vel, 7)) <= 2e-5
    # batch of 2 at 256^3: entries independent (size-independent property)
    vb = torch.cat([t(flow), t(flow).flip(0) * 0.5], 0).to(cuda)
    o = vxm.layers.VecInt(full, 3)(vb)
    assert torch.equal(o[:1], vxm.layers.VecInt(full, 3)(vb[:1]))


def test_config5_semisupervised_composition(vxm, cuda):
    shape = (32, 48, 40)
    half = tuple(s // 2 for s in shape)
    nlab = 30
    lab_m = cases.label_volume(311, shape, nlab)
    lab_f = cases.label_volume(312, shape, nlab)
    pos_flow = cases.smooth_field(313, 3, shape, scale=3.0)

    def onehot(lab):
        oh = (lab[:, 0, ::2, ::2, ::2][:, None] == np.arange(nlab, dtype=np.float32)[None, :, None, None, None])
        return oh.astype(np.float32)

    seg_m, seg_f = onehot(lab_m), onehot(lab_f)           # (1, 30, 16, 24, 20), generators.py:163-167 sub-sampling
    # --- oracle: ResizeTransform(2) -> linear warp of the prob-seg -> Dice, with autograd through the flow
    fc = t(pos_flow).double().requires_grad_(True)
    seg_flow_c = ref_torch.resize_transform(fc, 2)
    warped_c = ref_torch.spatial_transform(t(seg_m).double(), seg_flow_c)
    loss_c = ref_torch.dice_loss(t(seg_f).double(), warped_c)
    loss_c.backward()
    # --- B200 path
    fg = t(pos_flow).to(cuda).requires_grad_(True)
    seg_flow_g = vxm.layers.ResizeTransform(2, 3)(fg)
    warped_g = vxm.layers.SpatialTransformer(half)(t(seg_m).to(cuda), seg_flow_g)
    loss_g = vxm.losses.Dice().loss(t(seg_f).to(cuda), warped_g)
    loss_g.backward()
    assert rel(warped_g.detach().cpu().numpy(), warped_c.detach().numpy()) <= 1e-5
    assert abs(float(loss_g) - float(loss_c)) <= 1e-5 * abs(float(loss_c))
    assert rel(fg.grad.cpu().numpy(), fc.grad.numpy()) <= 1e-4
    # --- label-map propagation (evaluation): nearest-neighbour warp at full resolution, bit-exact
    moved = vxm.layers.SpatialTransformer(shape, mode="nearest")(t(lab_m).to(cuda), t(pos_flow).to(cuda)).cpu().numpy()
    assert np.array_equal(moved, spec_np.warp(lab_m, pos_flow, mode="nearest"))
    assert set(np.unique(moved)).issubset(set(np.unique(lab_m)) | {0.0})
