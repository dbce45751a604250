"""Tensor helpers (reference VoGE/Utils.py: ind_sel :13-31, ind_fill :34-56, Reshaper :59-76, Batchifier :79-176,
DataParallelBatchifier :179-333, rotation_theta :336-359, eye_like :9).  The renderer itself never calls the
Batchifiers: multi-GPU rendering shards views over torch.distributed ranks -- see voge_b200/distributed.py."""
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


# ---- chunked execution helpers (reference VoGE/Utils.py:59-333) --------------------------------------------------
# Not on the renderer's path (multi-GPU rendering shards views over torch.distributed ranks, voge_b200/distributed.py);
# provided so that `from VoGE.Utils import Batchifier` and the converters' large-cloud branch keep working.
class Reshaper(object):
    """Concatenates per-chunk results along `tar_index` and restores the flattened dims to `tar_shape`
    (reference :59-76).  Scalars (0-dim tensors / numbers) are summed."""

    def __init__(self, tar_shape, tar_index):
        self.tar_shape, self.tar_index = tuple(tar_shape), tar_index

    def __call__(self, x_):
        if isinstance(x_, (list, tuple)):
            if len(x_) == 0:
                return tuple()
            if isinstance(x_[0], (float, int)):
                return sum(x_)
            if torch.is_tensor(x_[0]) and x_[0].dim() == 0:
                return torch.sum(torch.stack(list(x_)))
            x_ = torch.cat(list(x_), dim=self.tar_index)
        if x_ is None:
            return None
        return x_.view(*self.tar_shape + tuple(x_.shape[self.tar_index + 1:]))


class Batchifier(object):
    """Decorator: run `func` on slices of `batch_size` along the flattened `target_dims` (or along everything but
    `remain_dims`) of the keyword arguments named in `batch_args`, and stitch the results back (reference :79-176).
    `tbar` is accepted for compatibility (the reference's callers pass it)."""

    def __init__(self, batch_size, batch_args, target_dims=None, remain_dims=None, tbar=False, **_ignored):
        self.batch_args = (batch_args,) if isinstance(batch_args, str) else tuple(batch_args)
        assert len(self.batch_args) > 0
        as_tuple = lambda d: None if d is None else ((d,) if isinstance(d, int) else tuple(d))
        self.target_dims, self.remain_dims = as_tuple(target_dims), (None if target_dims is not None else as_tuple(remain_dims))
        self.batch_size = int(batch_size)

    def _flatten(self, kwargs):
        """-> (kwargs with the batched tensors viewed with ONE flattened dim, index of that dim, shape to restore)"""
        kwargs = dict(kwargs)
        recorded, save_idx = None, None
        for k in self.batch_args:
            t = kwargs[k]
            assert isinstance(t, torch.Tensor)
            nd = t.dim()
            if self.target_dims is not None:
                flat = tuple(d % nd for d in self.target_dims)
            else:
                keep = tuple(d % nd for d in self.remain_dims)
                flat = tuple(i for i in range(nd) if i not in keep)
            assert flat == tuple(range(flat[0], flat[0] + len(flat))), 'the batched dims must be adjacent'
            shape = list(t.shape[:flat[0]]) + [-1] + list(t.shape[flat[-1] + 1:])
            restore = tuple(t.shape[:flat[-1] + 1])
            if recorded is None:
                recorded, save_idx = restore, flat[0]
            else:
                assert recorded == restore
            kwargs[k] = t.reshape(*shape)
        assert recorded is not None, 'No batchify parameters found!'
        return kwargs, save_idx, recorded

    def _slices(self, kwargs, save_idx, start, stop):
        sl = (slice(None),) * save_idx + (slice(start, stop),)
        return {k: (v[sl] if k in self.batch_args else v) for k, v in kwargs.items()}

    @staticmethod
    def _stitch(out, reshape):
        if isinstance(out[0], tuple):
            return tuple(reshape([o[i] for o in out]) for i in range(len(out[0])))
        return reshape(out)

    def __call__(self, func):
        def wrapper(*args, **kwargs):
            kw, save_idx, recorded = self._flatten(kwargs)
            total = kw[self.batch_args[0]].shape[save_idx]
            out = [func(*args, **self._slices(kw, save_idx, s, s + self.batch_size))
                   for s in range(0, max(total, 1), self.batch_size)]
            return self._stitch(out, Reshaper(recorded, save_idx))
        return wrapper


class DataParallelBatchifier(Batchifier):
    """Batchifier whose chunks are split over several GPUs, one Python thread per device (reference :179-333).
    Tensors among the keyword arguments are moved to the worker's device; results come back to the inputs' device."""

    def __init__(self, batch_size, batch_args, target_dims=None, remain_dims=None, device=None, **kw):
        super().__init__(batch_size, batch_args, target_dims=target_dims, remain_dims=remain_dims, **kw)
        self.device = [torch.device('cuda:%d' % i) for i in range(torch.cuda.device_count())] if device is None else list(device)
        self.n_gpus = len(self.device)

    def __call__(self, func):
        import threading

        def wrapper(*args, **kwargs):
            kw, save_idx, recorded = self._flatten(kwargs)
            home = kw[self.batch_args[0]].device
            total = kw[self.batch_args[0]].shape[save_idx]
            out = []
            for s in range(0, max(total, 1), self.batch_size):
                n = min(self.batch_size, total - s)
                per = (n - 1) // self.n_gpus + 1
                results, threads = {}, []

                def work(j, dev, sub):
                    try:
                        with torch.cuda.device(dev):
                            results[j] = func(*args, **sub)
                    except Exception as e:          # re-raised in the caller's thread
                        results[j] = e
                for j, dev in enumerate(self.device):
                    a, b = s + min(j * per, n), s + min((j + 1) * per, n)
                    if b <= a:
                        continue
                    sub = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in self._slices(kw, save_idx, a, b).items()}
                    threads.append(threading.Thread(target=work, args=(j, dev, sub)))
                for t in threads:
                    t.start()
                for t in threads:
                    t.join()
                for j in sorted(results):
                    r = results[j]
                    if isinstance(r, Exception):
                        raise r
                    out.append(r.to(home) if torch.is_tensor(r) else tuple(t.to(home) for t in r))
            return self._stitch(out, Reshaper(recorded, save_idx))
        return wrapper
