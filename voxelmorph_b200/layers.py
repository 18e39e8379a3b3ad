"""SpatialTransformer / VecInt / ResizeTransform with the reference's surface
(reference voxelmorph/torch/layers.py) over the sm_100a kernels of libvxm_b200.so.

Same class names, constructor arguments, forward signatures and error behaviour as the
reference; the arithmetic runs in hand-written CUDA kernels reached through the C ABI
(include/vxm_b200.h).  There is no CPU path.
"""
import math
import os

import torch
import torch.nn as nn

from . import _lib

MODE_LINEAR, MODE_NEAREST = 0, 1
ARITH_TRUE_DIV, ARITH_RECIPROCAL, ARITH_FAST = 0, 1, 2


def default_arith():
    """How `loc / (S-1)` (reference layers.py:37) is rounded.  'cpu' (default) replays torch's CPU
    true division — the oracle this repo is bit-exact against; 'cuda' replays torch's CUDA
    multiply-by-reciprocal."""
    return ARITH_RECIPROCAL if os.environ.get("VXM_B200_NEAREST_ARITH", "cpu") == "cuda" else ARITH_TRUE_DIV


def linear_arith():
    """Coordinate arithmetic of the LINEAR resampler (SpatialTransformer 'bilinear', VecInt).  Default 'fast':
    coord = (p + flow) * (Ssrc-1)/(S-1) — the reference's map without its fp32 normalise / un-normalise round trip;
    agrees with the replayed arithmetic to a few 1e-6 of the value range (north_star asks 1e-4) and lets the kernels run
    memory bound.  VXM_B200_LINEAR_ARITH=exact replays torch's arithmetic op for op (bit-identical to the torch CPU
    reference on every input tried), like the nearest mode always does."""
    return default_arith() if os.environ.get("VXM_B200_LINEAR_ARITH", "fast") == "exact" else ARITH_FAST


def _dims(t):
    """(B, C, D, H, W, nd) of a (B,C,[D,]H,W) tensor; 2-D is carried as D == 1."""
    if t.dim() == 5:
        B, C, D, H, W = t.shape
        return B, C, D, H, W, 3
    if t.dim() == 4:
        B, C, H, W = t.shape
        return B, C, 1, H, W, 2
    raise _lib.VxmError("voxelmorph_b200: expected a (B,C,H,W) or (B,C,D,H,W) tensor, got shape %s" % (tuple(t.shape),))


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, flow, mode, arith):
        _lib.require_cuda(src, flow, what="SpatialTransformer")
        src, flow = _lib.contig(src), _lib.contig(flow)
        B, C, Ds, Hs, Ws, nd = _dims(src)
        Bf, Cf, D, H, W, ndf = _dims(flow)
        if nd != ndf or Cf != nd or Bf != B:
            raise _lib.VxmError("SpatialTransformer: src %s and flow %s are inconsistent"
                                % (tuple(src.shape), tuple(flow.shape)))
        out = torch.empty((B, C) + tuple(flow.shape[2:]), dtype=torch.float32, device=src.device)
        lib = _lib.load()
        _lib.check(lib.vxm_warp_fwd(_lib.ptr(src), _lib.ptr(flow), _lib.ptr(out), B, C, Ds, Hs, Ws, D, H, W, nd,
                                    mode, arith, _lib.stream_ptr()), "vxm_warp_fwd")
        ctx.save_for_backward(src, flow)
        ctx.cfg = (mode, arith)
        return out

    @staticmethod
    def backward(ctx, gout):
        src, flow = ctx.saved_tensors
        mode, arith = ctx.cfg
        gout = _lib.contig(gout)
        B, C, Ds, Hs, Ws, nd = _dims(src)
        _, _, D, H, W, _ = _dims(flow)
        need_src, need_flow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gsrc = torch.zeros_like(src) if need_src else None
        gflow = None
        if need_flow:
            gflow = torch.zeros_like(flow) if mode == MODE_NEAREST else torch.empty_like(flow)
        if need_src or (need_flow and mode == MODE_LINEAR):
            lib = _lib.load()
            _lib.check(lib.vxm_warp_bwd(_lib.ptr(gout), _lib.ptr(src), _lib.ptr(flow), _lib.ptr(gsrc),
                                        _lib.ptr(gflow) if mode == MODE_LINEAR else None,
                                        B, C, Ds, Hs, Ws, D, H, W, nd, mode, arith, _lib.stream_ptr()),
                       "vxm_warp_bwd")
        return gsrc, gflow, None, None


class SpatialTransformer(nn.Module):
    """N-D spatial transformer (reference layers.py:6-48).

    `size` is kept for signature compatibility; no identity-grid buffer is materialised (the
    reference registers an 82.6 MB `grid` buffer per instance at 160x192x224) — checkpoints
    never contain it (modelio drops `*.grid`), so state_dict compatibility is unaffected.
    """

    def __init__(self, size, mode='bilinear'):
        super().__init__()
        self.mode = mode
        self.size = tuple(int(s) for s in size)

    def forward(self, src, flow):
        if self.mode == 'bilinear':
            m = MODE_LINEAR
        elif self.mode == 'nearest':
            m = MODE_NEAREST
        else:
            # F.grid_sample raises for anything but bilinear / nearest / bicubic; bicubic is 4-D only
            raise ValueError("nn.functional.grid_sample(): expected mode to be 'bilinear' or 'nearest', "
                             "but got: '%s'" % self.mode)
        return _WarpFn.apply(src, flow, m, linear_arith() if m == MODE_LINEAR else default_arith())


class _VecIntFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, nsteps, arith):
        _lib.require_cuda(vec, what="VecInt")
        vec = _lib.contig(vec)
        B, C, D, H, W, nd = _dims(vec)
        if C != nd:
            raise _lib.VxmError("VecInt: expected %d flow channels, got %d" % (nd, C))
        lib = _lib.load()
        out = torch.empty_like(vec)
        need_grad = ctx.needs_input_grad[0]
        if arith == ARITH_FAST and (nd != 3 or nsteps < 1 or min(D, H, W) < 2):
            arith = default_arith()      # the float4 fast path is 3-D only
        states = work = None
        if nsteps > 0:
            if arith == ARITH_FAST:
                if need_grad:
                    states = torch.empty(int(lib.vxm_vecint_fast_states_bytes(B, D, H, W, nsteps)), dtype=torch.uint8, device=vec.device)
                else:
                    work = torch.empty(int(lib.vxm_vecint_fast_work_bytes(B, D, H, W, 0)), dtype=torch.uint8, device=vec.device)
            elif need_grad:
                states = torch.empty((nsteps,) + tuple(vec.shape), dtype=torch.float32, device=vec.device)
            else:
                work = torch.empty_like(vec)
        _lib.check(lib.vxm_vecint_fwd(_lib.ptr(vec), _lib.ptr(out), _lib.ptr(states), _lib.ptr(work), B, D, H, W, nd,
                                      nsteps, arith, _lib.stream_ptr()), "vxm_vecint_fwd")
        ctx.states = states
        ctx.cfg = (nsteps, arith, (B, D, H, W, nd))
        return out

    @staticmethod
    def backward(ctx, gout):
        nsteps, arith, (B, D, H, W, nd) = ctx.cfg
        gout = _lib.contig(gout)
        gvel = torch.empty_like(gout)
        lib = _lib.load()
        work = None
        if nsteps > 0:
            if arith == ARITH_FAST:
                work = torch.empty(int(lib.vxm_vecint_fast_work_bytes(B, D, H, W, 1)), dtype=torch.uint8, device=gout.device)
            else:
                work = torch.empty((2,) + tuple(gout.shape), dtype=torch.float32, device=gout.device)
        _lib.check(lib.vxm_vecint_bwd(_lib.ptr(gout), _lib.ptr(ctx.states), _lib.ptr(gvel), _lib.ptr(work), B, D, H, W,
                                      nd, nsteps, arith, _lib.stream_ptr()), "vxm_vecint_bwd")
        return gvel, None, None


class VecInt(nn.Module):
    """Integrates a vector field via scaling and squaring (reference layers.py:51-68), all
    `nsteps` squarings fused into one cooperative kernel launch."""

    def __init__(self, inshape, nsteps):
        super().__init__()
        assert nsteps >= 0, 'nsteps should be >= 0, found: %d' % nsteps
        self.nsteps = nsteps
        self.scale = 1.0 / (2 ** self.nsteps)
        self.transformer = SpatialTransformer(inshape)

    def forward(self, vec):
        return _VecIntFn.apply(vec, self.nsteps, linear_arith())


class _ResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, out_spatial, pre, post):
        _lib.require_cuda(x, what="ResizeTransform")
        x = _lib.contig(x)
        B, C, Di, Hi, Wi, nd = _dims(x)
        if nd == 3:
            Do, Ho, Wo = out_spatial
        else:
            Do, (Ho, Wo) = 1, out_spatial
        out = torch.empty((B, C) + tuple(out_spatial), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        _lib.check(lib.vxm_resize_fwd(_lib.ptr(x), _lib.ptr(out), B, C, Di, Hi, Wi, Do, Ho, Wo, pre, post,
                                      _lib.stream_ptr()), "vxm_resize_fwd")
        ctx.cfg = (B, C, Di, Hi, Wi, Do, Ho, Wo, pre, post, tuple(x.shape))
        return out

    @staticmethod
    def backward(ctx, gout):
        B, C, Di, Hi, Wi, Do, Ho, Wo, pre, post, xshape = ctx.cfg
        gout = _lib.contig(gout)
        gx = torch.empty(xshape, dtype=torch.float32, device=gout.device)
        lib = _lib.load()
        _lib.check(lib.vxm_resize_bwd(_lib.ptr(gout), _lib.ptr(gx), B, C, Di, Hi, Wi, Do, Ho, Wo, pre, post,
                                      _lib.stream_ptr()), "vxm_resize_bwd")
        return gx, None, None, None


class ResizeTransform(nn.Module):
    """Resize a transform: resize the vector field *and* rescale it (reference layers.py:71-97)."""

    def __init__(self, vel_resize, ndims):
        super().__init__()
        self.factor = 1.0 / vel_resize
        self.mode = 'linear'
        if ndims == 2:
            self.mode = 'bi' + self.mode
        elif ndims == 3:
            self.mode = 'tri' + self.mode

    def forward(self, x):
        if self.factor == 1:
            return x  # layers.py:96: "don't do anything if resize is 1"
        nd = x.dim() - 2
        if (nd == 2 and self.mode != 'bilinear') or (nd == 3 and self.mode != 'trilinear') or nd not in (2, 3):
            raise NotImplementedError("Got %dD input, but interpolation mode '%s' needs a matching dimensionality"
                                      % (x.dim(), self.mode))
        # F.interpolate(scale_factor=...) output size: floor(in * scale)
        out_spatial = tuple(int(math.floor(float(s) * self.factor)) for s in x.shape[2:])
        if self.factor < 1:   # resize first, then rescale (layers.py:86-89)
            return _ResizeFn.apply(x, out_spatial, 1.0, float(self.factor))
        return _ResizeFn.apply(x, out_spatial, float(self.factor), 1.0)  # layers.py:91-94
