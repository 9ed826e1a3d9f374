"""world_size-2 gloo checks (CPU) of the multi-GPU host logic: the flat gradient all-reduce of the data-parallel
training step and the contiguous window sharding of sliding-window inference."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _grad_case(rank, world):
    from hybrid_ctunet_b200.dp import GradientAllReduce
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((3, 5), float(rank + 1))
    params[1].grad = None                      # never-used weight: must stay None on every rank
    params[2].grad = torch.arange(4.0).reshape(2, 2) * (rank + 1)
    n = GradientAllReduce(params).reduce()
    return n, params[0].grad.clone(), params[1].grad, params[2].grad.clone()


def test_gradient_allreduce_is_mean_and_skips_none():
    out = _run(_grad_case)
    for rank in (0, 1):
        n, g0, g1, g2 = out[rank]
        assert n == 19 and g1 is None
        assert torch.equal(g0, torch.full((3, 5), 1.5))
        assert torch.equal(g2, torch.arange(4.0).reshape(2, 2) * 1.5)


def _flat_case(rank, world):
    """Gradients that are aligned slices of ONE flat buffer (how the engine hands them out) are reduced in place."""
    from hybrid_ctunet_b200.dp import GradientAllReduce
    params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    flat = torch.full((16 + 8 + 4,), float("nan"))          # gaps between the 16-byte-aligned slices hold garbage
    views = [flat[0:15].view(3, 5), flat[16:23].view(7), flat[24:28].view(2, 2)]
    for v, p in zip(views, params):
        v.fill_(float(rank + 1))
        p.grad = v
    ar = GradientAllReduce(params)
    n = ar.reduce()
    same_storage = all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in params)
    return n, [p.grad.clone() for p in params], same_storage, ar._flat is None


def test_gradient_allreduce_in_place_on_one_flat_buffer():
    out = _run(_flat_case)
    for rank in (0, 1):
        n, gs, same, no_copy = out[rank]
        assert n == 26 and same and no_copy
        for g in gs:
            assert torch.equal(g, torch.full_like(g, 1.5))


def _shard_case(rank, world):
    """The window ranges the sharded sliding window assigns (hybrid_ctunet_b200/sliding_window.py) tile the window list
    exactly once, contiguously."""
    from hybrid_ctunet_b200.sliding_window import dense_patch_starts, get_scan_interval
    image, roi = (512, 512, 256), (96, 96, 96)
    starts = dense_patch_starts(image, roi, get_scan_interval(image, roi, 3, 0.5))
    total = len(starts)
    per = -(-total // world)
    lo, hi = min(rank * per, total), min((rank + 1) * per, total)
    cover = torch.zeros(total)
    cover[lo:hi] = 1
    dist.all_reduce(cover)
    return total, lo, hi, bool((cover == 1).all())


def test_window_sharding_covers_every_window_once():
    out = _run(_shard_case)
    assert out[0][0] == 500 and out[0][3] and out[1][3]
    assert out[0][1:3] == (0, 250) and out[1][1:3] == (250, 500)
