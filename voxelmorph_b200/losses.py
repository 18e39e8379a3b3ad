"""NCC / MSE / Dice / Grad losses with the reference's surface
(reference voxelmorph/torch/losses.py): plain classes whose bound `.loss(y_true, y_pred)`
returns a 0-d tensor supporting `.item()`, `*`, `+`, `.backward()`.
Each loss is one fused sm_100a kernel (plus a fused backward) from libvxm_b200.so.
"""
import numpy as np
import torch

from . import _lib
from .layers import _dims


def _scalar(device):
    return torch.empty((), dtype=torch.float32, device=device)


class _NccFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_true, y_pred, win):
        _lib.require_cuda(y_true, y_pred, what="NCC")
        I, J = _lib.contig(y_true), _lib.contig(y_pred)
        B, C, D, H, W, nd = _dims(I)
        if C != 1 or tuple(J.shape) != tuple(I.shape):
            # the reference's ones(1,1,*win) filter only accepts single-channel volumes (losses.py:29,51)
            raise _lib.VxmError("NCC: expected two single-channel volumes of equal shape, got %s and %s"
                                % (tuple(I.shape), tuple(J.shape)))
        wd, wh, ww = (1, win[0], win[1]) if nd == 2 else tuple(win)
        lib = _lib.load()
        loss = _scalar(I.device)
        need = ctx.needs_input_grad[1]
        if ctx.needs_input_grad[0]:
            raise _lib.VxmError("NCC: gradients w.r.t. y_true are not implemented (the training loop "
                                "only differentiates y_pred, reference scripts/torch/train.py:210)")
        saved = torch.empty((B, 3, D, H, W), dtype=torch.float32, device=I.device) if need else None
        ws = _lib.reduce_workspace(I.device)
        _lib.check(lib.vxm_ncc_fwd(_lib.ptr(I), _lib.ptr(J), _lib.ptr(loss), _lib.ptr(saved), _lib.ptr(ws),
                                   B, D, H, W, wd, wh, ww, _lib.stream_ptr()), "vxm_ncc_fwd")
        ctx.save_for_backward(I, J)
        ctx.saved_fields = saved
        ctx.cfg = (B, D, H, W, wd, wh, ww)
        return loss

    @staticmethod
    def backward(ctx, gl):
        I, J = ctx.saved_tensors
        B, D, H, W, wd, wh, ww = ctx.cfg
        gl = gl.contiguous().float()
        gJ = torch.empty_like(J)
        lib = _lib.load()
        _lib.check(lib.vxm_ncc_bwd(_lib.ptr(I), _lib.ptr(J), _lib.ptr(ctx.saved_fields), _lib.ptr(gl), _lib.ptr(gJ),
                                   B, D, H, W, wd, wh, ww, _lib.stream_ptr()), "vxm_ncc_bwd")
        return None, gJ, None


class NCC:
    """Local (over window) normalized cross correlation loss (reference losses.py:7-67)."""

    def __init__(self, win=None):
        self.win = win

    def loss(self, y_true, y_pred):
        ndims = len(list(y_true.size())) - 2
        assert ndims in [1, 2, 3], "volumes should be 1 to 3 dimensions. found: %d" % ndims
        if ndims == 1:
            raise _lib.VxmError("NCC: 1-D volumes are not supported by the B200 path")
        win = [9] * ndims if self.win is None else list(self.win)
        return _NccFn.apply(y_true, y_pred, tuple(int(w) for w in win))


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_true, y_pred):
        _lib.require_cuda(y_true, y_pred, what="MSE")
        if tuple(y_true.shape) != tuple(y_pred.shape):
            y_true, y_pred = torch.broadcast_tensors(y_true, y_pred)
        a, b = _lib.contig(y_true), _lib.contig(y_pred)
        lib = _lib.load()
        loss = _scalar(a.device)
        _lib.check(lib.vxm_mse_fwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(loss), _lib.ptr(_lib.reduce_workspace(a.device)),
                                   a.numel(), _lib.stream_ptr()), "vxm_mse_fwd")
        ctx.save_for_backward(a, b)
        return loss

    @staticmethod
    def backward(ctx, gl):
        a, b = ctx.saved_tensors
        gl = gl.contiguous().float()
        gp = torch.empty_like(b)
        lib = _lib.load()
        _lib.check(lib.vxm_mse_bwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(gl), _lib.ptr(gp), a.numel(), _lib.stream_ptr()),
                   "vxm_mse_bwd")
        gt = -gp if ctx.needs_input_grad[0] else None
        return gt, gp


class MSE:
    """Mean squared error loss (reference losses.py:70-76)."""

    def loss(self, y_true, y_pred):
        return _MseFn.apply(y_true, y_pred)


class _DiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_true, y_pred):
        _lib.require_cuda(y_true, y_pred, what="Dice")
        a, b = _lib.contig(y_true), _lib.contig(y_pred)
        if tuple(a.shape) != tuple(b.shape) or a.dim() < 3:
            raise _lib.VxmError("Dice: expected equal (B,L,*vol) shapes, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
        BL = a.shape[0] * a.shape[1]
        V = a.numel() // BL
        lib = _lib.load()
        loss = _scalar(a.device)
        sums = torch.empty((BL, 2), dtype=torch.float32, device=a.device)
        work = torch.empty(int(lib.vxm_dice_workspace_bytes(BL)), dtype=torch.uint8, device=a.device)
        _lib.check(lib.vxm_dice_fwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(loss), _lib.ptr(sums), _lib.ptr(work), BL, V,
                                    _lib.stream_ptr()), "vxm_dice_fwd")
        ctx.save_for_backward(a, sums)
        ctx.cfg = (BL, V)
        if ctx.needs_input_grad[0]:
            raise _lib.VxmError("Dice: gradients w.r.t. y_true are not implemented")
        return loss

    @staticmethod
    def backward(ctx, gl):
        a, sums = ctx.saved_tensors
        BL, V = ctx.cfg
        gl = gl.contiguous().float()
        gp = torch.empty_like(a)
        lib = _lib.load()
        _lib.check(lib.vxm_dice_bwd(_lib.ptr(a), _lib.ptr(sums), _lib.ptr(gl), _lib.ptr(gp), BL, V, _lib.stream_ptr()),
                   "vxm_dice_bwd")
        return None, gp


class Dice:
    """N-D dice for segmentation (reference losses.py:79-90)."""

    def loss(self, y_true, y_pred):
        return _DiceFn.apply(y_true, y_pred)


class _GradFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, penalty, mult):
        _lib.require_cuda(y, what="Grad")
        y = _lib.contig(y)
        B, C, D, H, W, nd = _dims(y)
        lib = _lib.load()
        loss = _scalar(y.device)
        _lib.check(lib.vxm_gradloss_fwd(_lib.ptr(y), _lib.ptr(loss), _lib.ptr(_lib.reduce_workspace(y.device)), B, C, D,
                                        H, W, nd, penalty, mult, _lib.stream_ptr()), "vxm_gradloss_fwd")
        ctx.save_for_backward(y)
        ctx.cfg = (B, C, D, H, W, nd, penalty, mult)
        return loss

    @staticmethod
    def backward(ctx, gl):
        (y,) = ctx.saved_tensors
        B, C, D, H, W, nd, penalty, mult = ctx.cfg
        gl = gl.contiguous().float()
        gy = torch.empty_like(y)
        lib = _lib.load()
        _lib.check(lib.vxm_gradloss_bwd(_lib.ptr(y), _lib.ptr(gl), _lib.ptr(gy), B, C, D, H, W, nd, penalty, mult,
                                        _lib.stream_ptr()), "vxm_gradloss_bwd")
        return gy, None, None


class Grad:
    """N-D gradient loss (reference losses.py:93-135)."""

    def __init__(self, penalty='l1', loss_mult=None):
        self.penalty = penalty
        self.loss_mult = loss_mult

    def loss(self, _, y_pred):
        if self.penalty == 'l1':
            p = 1
        else:
            assert self.penalty == 'l2', 'penalty can only be l1 or l2. Got: %s' % self.penalty
            p = 2
        mult = 1.0 if self.loss_mult is None else float(self.loss_mult)
        return _GradFn.apply(y_pred, p, mult)
