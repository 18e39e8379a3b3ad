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
