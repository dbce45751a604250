"""Multi-GPU layer: one process per GPU (torchrun), views sharded by camera, one all-reduce of the
Gaussian-parameter gradients per optimisation step over NCCL (NVLink 5 / NVSwitch).

The reference has no distributed code; its only parallel helper is the thread-per-GPU
DataParallelBatchifier (reference VoGE/Utils.py:179-333), which the renderer never calls.  Rays of
different views are independent, every rank holds the full Gaussian set (<= 60 MB at N = 1M), so the
path shards by view with NO data-path collective; the only exchange is the gradient all-reduce of a
fitting loop.
"""
import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank).  Single-process runs need no process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_views(n_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of views [first, first+count) owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_views, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def allreduce_gradients(params: Sequence[torch.Tensor], average: bool = False):
    """Sum (or average) the .grad of `params` over all ranks with ONE all-reduce on a flat fp32
    bucket [d verts ; d sigmas ; d colours ...].  Parameters without a gradient contribute zeros so
    that every rank issues the same collective."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    grads = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
