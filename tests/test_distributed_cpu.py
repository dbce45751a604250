"""World-size-2 gloo test (CPU) of the multi-GPU host logic: view sharding covers every view exactly
once and the flat-bucket gradient all-reduce equals the single-process sum."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voge_b200.distributed import shard_view_indices, shard_views


def test_shard_view_indices_round_robin():
    for n in (1, 5, 8, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [shard_view_indices(n, r, world) for r in range(world)]
            assert sorted(v for p in parts for v in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard_view_indices(64, 3, 8) == [3, 11, 19, 27, 35, 43, 51, 59]


def test_shard_views_partitions():
    for n in (1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, count = shard_views(n, r, world)
                seen += list(range(first, first + count))
            assert seen == list(range(n))
    assert shard_views(64, 3, 8) == (24, 8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from voge_b200.distributed import allreduce_gradients, barrier, init_from_env, max_over_ranks
    r, w, _ = init_from_env("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(0)
    verts = torch.nn.Parameter(torch.rand(50, 3, generator=g))
    sig = torch.nn.Parameter(torch.rand(50, 3, 3, generator=g))
    col = torch.nn.Parameter(torch.rand(50, 3, generator=g))
    first, count = shard_views(6, rank, world)
    # per-view pseudo-gradients: view v contributes (v+1) * param
    verts.grad = sum((v + 1) * verts.detach() for v in range(first, first + count))
    sig.grad = sum((v + 1) * sig.detach() for v in range(first, first + count))
    col.grad = None                       # a parameter without gradient on this rank must still join
    allreduce_gradients([verts, sig, col])
    total = sum(range(1, 7))
    ok = torch.allclose(verts.grad, total * verts.detach()) and torch.allclose(sig.grad, total * sig.detach())
    ok = ok and float(col.grad.abs().sum()) == 0.0
    ok = ok and max_over_ranks(float(rank)) == float(world - 1)
    # persistent bucket: .grad tensors alias one flat buffer, autograd accumulates in place over several
    # backward calls (chunks of views), one in-place all-reduce, zero() between steps
    from voge_b200.distributed import GradientBucket
    a = torch.nn.Parameter(torch.ones(4, 3))
    b = torch.nn.Parameter(torch.ones(5))
    bucket = GradientBucket([a, b])
    for step in range(2):
        bucket.zero()
        for v in range(first, first + count):          # one backward per view of this rank
            ((v + 1) * (a.sum() + 2 * b.sum())).backward()
        ok = ok and bucket.attached() and a.grad.data_ptr() == bucket.flat.data_ptr()
        bucket.allreduce()
        ok = ok and torch.allclose(a.grad, torch.full((4, 3), float(total))) and torch.allclose(b.grad, torch.full((5,), 2.0 * total))
    bucket.allreduce(average=True)
    ok = ok and torch.allclose(a.grad, torch.full((4, 3), float(total)))      # the average of equal buffers
    a.grad = None
    ok = ok and not bucket.attached()
    barrier()
    out[rank] = ok
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)) and len(out) == world
