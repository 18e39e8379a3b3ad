"""Flat-buffer fused Adam (replaces torch.optim.Adam of reference scripts/torch/train.py:161,220).

All parameters are re-pointed into ONE contiguous fp32 buffer (and all gradients into another),
so that the optimizer step is a single kernel launch and the data-parallel gradient exchange is
a single NCCL allreduce over one buffer (voxelmorph_b200/dist.py).  `state_dict` keys and
parameter shapes are untouched: `param.data` / `param.grad` become views.
"""
import torch

from . import _lib


class FlatParams:
    """Flattens a module's parameters (and their .grad) into two contiguous fp32 buffers."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatParams: no trainable parameters")
        dev = self.params[0].device
        _lib.require_cuda(*self.params, what="FlatParams")
        n = sum(p.numel() for p in self.params)
        self.numel = n
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.grad[off:off + k].view_as(p)
                p._vxm_flat_grad = True     # engine_bf16 accumulates weight gradients straight into this view
                off += k

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p in self.params:  # re-attach views if autograd replaced .grad
            k = p.numel()
            g = self.grad[off:off + k].view_as(p)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g
            off += k


class FusedAdam:
    """Adam with torch.optim.Adam's update rule (bias-corrected, eps outside the sqrt), one launch."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, world_size=1):
        self.fp = params if isinstance(params, FlatParams) else FlatParams(list(params))
        self._group = dict(lr=lr, params=self.fp.params)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.m = torch.zeros_like(self.fp.flat)
        self.v = torch.zeros_like(self.fp.flat)
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.fp.flat.device)   # device-side step count (graph capture)
        self.grad_scale = 1.0 / world_size

    @property
    def lr(self):
        return self._group["lr"]

    @lr.setter
    def lr(self, value):
        self._group["lr"] = value

    def zero_grad(self, set_to_none=False):
        self.fp.zero_grad()

    # torch.optim-like surface (LR schedulers read / write param_groups[0]['lr']; resume needs state_dict)
    @property
    def param_groups(self):
        return [self._group]

    def state_dict(self):
        return dict(step=int(self.step_dev.item()), m=self.m.clone(), v=self.v.clone(), lr=self._group["lr"], betas=self.betas,
                    eps=self.eps, weight_decay=self.weight_decay)

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.step_dev.fill_(int(sd["step"]))
        self.step_count = int(sd["step"])
        self._group["lr"] = sd["lr"]
        self.betas, self.eps, self.weight_decay = tuple(sd["betas"]), sd["eps"], sd["weight_decay"]

    def snapshot(self):
        """Parameters + optimizer state (device copies) — see trainer.GraphedTrainStep.capture."""
        return dict(flat=self.fp.flat.clone(), m=self.m.clone(), v=self.v.clone(), step_dev=self.step_dev.clone(), step=self.step_count)

    def restore(self, snap):
        with torch.no_grad():
            self.fp.flat.copy_(snap["flat"])
            self.m.copy_(snap["m"])
            self.v.copy_(snap["v"])
            self.step_dev.copy_(snap["step_dev"])
        self.step_count = snap["step"]
        from . import engine_bf16
        engine_bf16.bump_weights_epoch()

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        off = 0
        for p in self.fp.params:  # gradients produced outside the flat buffer are folded in
            k = p.numel()
            g = self.fp.grad[off:off + k]
            if p.grad is None:
                # model.zero_grad(set_to_none=True) or a parameter that received no gradient: its slot must not replay
                # the previous step's gradient (torch.optim.Adam skips such parameters; a zero gradient is the closest
                # the one-launch kernel gets — the moments decay, the weight barely moves)
                g.zero_()
                p.grad = g.view_as(p)
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad.reshape(-1))
                p.grad = g.view_as(p)
            off += k
        lib = _lib.load()
        _lib.check(lib.vxm_adam_step_dev(_lib.ptr(self.fp.flat), _lib.ptr(self.fp.grad), _lib.ptr(self.m), _lib.ptr(self.v),
                                         self.fp.numel, _lib.ptr(self.step_dev), float(self._group['lr']), self.betas[0], self.betas[1], self.eps,
                                         self.weight_decay, self.grad_scale, _lib.stream_ptr()), "vxm_adam_step_dev")
        from . import engine_bf16
        engine_bf16.bump_weights_epoch()   # parameters changed behind torch's version counter
