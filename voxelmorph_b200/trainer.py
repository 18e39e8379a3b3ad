"""Whole-step CUDA-graph capture of the VxmDense training step.

The reference's loop (scripts/torch/train.py:199-220: forward, losses, backward, Adam) issues ~100 kernels per step
here; at ~10 ms per step the launch gaps are worth a millisecond.  `GraphedTrainStep` captures ONE full step
(zero-grad, forward, NCC/MSE + Grad losses, backward, the single gradient allreduce, fused Adam) into a CUDA graph on
static input buffers and replays it; nothing in the captured region reads host state (the Adam step count lives on
the device, packed bf16 weights are re-derived inside the graph).
"""
import torch

from . import dist as vdist
from . import losses


class GraphedTrainStep:
    def __init__(self, model, optimizer, image_loss="ncc", lam=0.01, int_downsize=2, warmup=3):
        self.model, self.opt = model, optimizer
        self.img = losses.NCC().loss if image_loss == "ncc" else losses.MSE().loss
        self.grad = losses.Grad("l2", loss_mult=int_downsize).loss
        self.lam = lam
        self.warmup = warmup
        self.graph = None
        self.src = self.trg = self.loss = None

    def _step(self):
        self.opt.zero_grad()
        y, flow = self.model(self.src, self.trg)
        loss = self.img(self.trg, y) + self.lam * self.grad(None, flow)
        loss.backward()
        vdist.allreduce_grads(self.opt.fp.grad)
        self.opt.step()
        return loss.detach()

    def capture(self, src, trg):
        """Warm up eagerly on a side stream, then capture one step.  `src` / `trg` fix the static shapes."""
        self.src, self.trg = torch.empty_like(src), torch.empty_like(trg)
        self.src.copy_(src)
        self.trg.copy_(trg)
        side = torch.cuda.Stream(device=src.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        return self

    def __call__(self, src=None, trg=None):
        """Replay.  If src/trg are given they are copied into the static buffers first (device or pinned host tensors)."""
        if src is not None:
            self.src.copy_(src, non_blocking=True)
            self.trg.copy_(trg, non_blocking=True)
        self.graph.replay()
        return self.loss
