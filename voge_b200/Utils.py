"""Tensor helpers on the hot path (reference VoGE/Utils.py: ind_sel :13-31, ind_fill :34-56,
rotation_theta :336-359, eye_like :9).  The reference's Batchifier / DataParallelBatchifier
(:80-333, thread-per-GPU scatter/gather, never called by the renderer) are replaced by camera
sharding over torch.distributed -- see voge_b200/distributed.py."""
import torch


def eye_like(tensor: torch.Tensor) -> torch.Tensor:
    n = tensor.shape[-1]
    eye = torch.eye(n, device=tensor.device, dtype=tensor.dtype)
    return eye.expand(tensor.shape[:-2] + (n, n))


def _broadcast_index(target: torch.Tensor, ind: torch.Tensor, dim: int):
    if ind.dim() <= dim:
        raise AssertionError("Index must have the target dim, but get dim: %d, ind shape: %s" % (dim, str(ind.shape)))
    lead = [ind.shape[k] if target.shape[k] == 1 else -1 for k in range(dim)]
    target = target.expand(*lead, *([-1] * (target.dim() - dim)))
    idx = ind
    trailing = target.shape[dim + 1:]
    if len(trailing):
        idx = idx.reshape(idx.shape + (1,) * len(trailing)).expand(*([-1] * (dim + 1)), *trailing)
    return target, idx


def ind_sel(target: torch.Tensor, ind: torch.Tensor, dim: int = 1) -> torch.Tensor:
    """gather along `dim` with the index broadcast over trailing dims:
    target [..., n, ...], ind [..., M] -> [..., M, ...]"""
    target, idx = _broadcast_index(target, ind, dim)
    return torch.gather(target, dim=dim, index=idx)


def ind_fill(target: torch.Tensor, ind: torch.Tensor, src, dim: int = 1) -> torch.Tensor:
    """scatter `src` (tensor or scalar) along `dim`, out of place."""
    target, idx = _broadcast_index(target, ind, dim)
    if torch.is_tensor(src):
        return target.scatter(dim=dim, index=idx, src=src)
    return target.scatter(dim=dim, index=idx, value=src)


def rotation_theta(theta, device_=None) -> torch.Tensor:
    """In-plane rotation matrices [[cos,-sin,0],[sin,cos,0],[0,0,1]] of shape (n,3,3)."""
    if isinstance(theta, float):
        theta = torch.full((1,), theta, device=device_ or "cpu")
    elif device_ is None:
        device_ = theta.device
    theta = theta.reshape(-1).to(device_)
    c, s = torch.cos(theta), torch.sin(theta)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([c, -s, z, s, c, z, z, z, o], dim=1).view(-1, 3, 3)
