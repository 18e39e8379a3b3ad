"""One process per GPU for an UNMODIFIED training script.

    python -m voxelmorph_b200.launch --nproc 8 scripts/torch/train.py --img-list list.txt --batch-size 1 ...

starts `nproc` copies of the script with torchrun's environment (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR=127.0.0.1,
MASTER_PORT) and appends `--gpu <local rank>` to each (the reference's train.py selects its device with --gpu and sets
CUDA_VISIBLE_DEVICES itself, scripts/torch/train.py:58,125).  Inside, every VxmDense turns data parallel by itself
(voxelmorph_b200.dist.TransparentDP): one flat-gradient allreduce per step, rank-0-only checkpoints.  `torchrun` works
as well when the script needs no per-rank argument.
"""
import argparse
import os
import socket
import subprocess
import sys


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m voxelmorph_b200.launch")
    ap.add_argument("--nproc", type=int, required=True, help="processes (= GPUs) on this node")
    ap.add_argument("--gpu-flag", default="--gpu", help="per-rank device flag appended to the script's arguments ('' to append nothing)")
    ap.add_argument("--master-port", type=int, default=0)
    ap.add_argument("script")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    port = a.master_port or _free_port()
    procs = []
    for r in range(a.nproc):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(a.nproc), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        cmd = [sys.executable, a.script] + list(a.args) + ([a.gpu_flag, str(r)] if a.gpu_flag else [])
        procs.append(subprocess.Popen(cmd, env=env))
    rc = 0
    for p in procs:
        rc = p.wait() or rc
    return rc


if __name__ == "__main__":
    sys.exit(main())
