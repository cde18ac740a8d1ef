"""Multi-GPU plumbing: frames are independent units, so ranks take disjoint frame sets and never
exchange point, grid or feature data; only the (3 + 3C) int64 IoU counts are all-reduced
(SURVEY.md section 8(e)).  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import os

import torch


def shard_frames(n_frames: int, rank: int, world_size: int, contiguous: bool = True):
    """Frame indices owned by ``rank``.  contiguous blocks (default) or round-robin ``f % world == rank``."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    if contiguous:
        per, rem = divmod(n_frames, world_size)
        lo = rank * per + min(rank, rem)
        return list(range(lo, lo + per + (1 if rank < rem else 0)))
    return list(range(rank, n_frames, world_size))


def init_distributed(backend: str | None = None):
    """Initialise from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world, device)."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if use_cuda else {}
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world, **kw)
    return rank, world, device
