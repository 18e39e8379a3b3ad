"""Data-parallel plumbing: one process per GPU, one volume pair per GPU per step, and exactly
one allreduce of the flat gradient buffer per step (replaces the reference's single-process
torch.nn.DataParallel, scripts/torch/train.py:151-154).  Works with the `nccl` backend on GPUs
and with `gloo` on CPU (used by the world_size-2 host-logic tests).
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; no-op for a single process."""
    world, rank, local = env_world()
    if world == 1:
        return world, rank, local
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return world, rank, local


def shard_indices(n_items, world, rank):
    """Rank r owns items r, r+world, r+2*world, ... (independent volume pairs; no data-path collective)."""
    return list(range(rank, n_items, world))


def broadcast_params(flat, src=0):
    """One-time broadcast of the flat parameter buffer from rank `src`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)


def allreduce_grads(flat_grad):
    """The step's single collective: SUM over ranks of the flat fp32 gradient buffer (1.31 MB for the
    default 3-D U-Net).  The 1/world factor is folded into the fused Adam (`grad_scale`)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)


def max_over_ranks(value, device):
    """max over ranks of a python float (timing)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------------
# Transparent data parallelism for the UNMODIFIED training loop (reference scripts/torch/train.py:143-161,199-220).
#
# The reference's own multi-GPU path is single-process nn.DataParallel (train.py:151-154).  Here N copies of the script
# are started, one per GPU (torchrun, or `python -m voxelmorph_b200.launch`); each builds its VxmDense as usual.  On its
# first forward a VxmDense that finds RANK / WORLD_SIZE in the environment
#   * joins the process group (NCCL on GPUs, gloo on CPU tensors),
#   * re-points its parameters into ONE flat fp32 buffer and broadcasts it from rank 0 (so every replica starts from
#     the same weights whatever each process drew at construction),
#   * hooks its parameters so that, once per backward, all gradients are averaged with ONE allreduce of one flat buffer
#     (queued as an autograd engine callback: it runs after the last gradient of the step has been accumulated),
#   * and `save()` writes the checkpoint on rank 0 only.
# torch.optim.Adam (train.py:161) works unchanged on the re-pointed parameters.  Nothing happens when WORLD_SIZE is
# absent or 1.
# ---------------------------------------------------------------------------------------------------------------------
class TransparentDP:
    def __init__(self, module, backend=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("TransparentDP: module has no trainable parameters")
        dev = self.params[0].device
        if backend is None:
            backend = "nccl" if dev.type == "cuda" else "gloo"
        world, rank, _ = env_world()
        if not dist.is_initialized():
            if backend == "nccl":
                torch.cuda.set_device(dev)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=self.params[0].dtype, device=dev)
        self.flat_grad = torch.zeros(n, dtype=self.params[0].dtype, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                off += k
        dist.broadcast(self.flat, src=0)
        self._queued = False
        self.allreduces = 0
        self._handles = [p.register_hook(self._on_grad) for p in self.params]

    # a gradient of this backward pass has been produced: make sure the finaliser is queued exactly once
    def _on_grad(self, grad):
        self.schedule()
        return grad

    def schedule(self):
        """Queue the gradient exchange for the end of the running backward pass (idempotent per pass)."""
        if not self._queued:
            self._queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

    @torch.no_grad()
    def _finalize(self):
        self._queued = False
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None:
                self.flat_grad[off:off + k].zero_()
            elif p.grad.data_ptr() != self.flat_grad[off:off + k].data_ptr():
                self.flat_grad[off:off + k].copy_(p.grad.reshape(-1))
            off += k
        dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
        self.flat_grad.mul_(1.0 / self.world)
        self.allreduces += 1
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad[off:off + k].data_ptr():
                p.grad.copy_(self.flat_grad[off:off + k].view_as(p))
            off += k

    def is_writer(self):
        return self.rank == 0


def attach_if_distributed(module):
    """TransparentDP for `module` when the process was started by torchrun (WORLD_SIZE > 1), else None.
    VXM_B200_TRANSPARENT_DP=0 disables it (e.g. when the caller drives `allreduce_grads` itself, like bench.py)."""
    world, _, _ = env_world()
    if world <= 1 or os.environ.get("VXM_B200_TRANSPARENT_DP", "1") == "0":
        return None
    return TransparentDP(module)
