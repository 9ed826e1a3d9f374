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


class _SubsetSGD(torch.optim.SGD):
    """SGD with the `only=` / `key=` arguments of hybrid_ctunet_b200.optim.AdamW.step (CPU stand-in)."""

    @torch.no_grad()
    def step(self, closure=None, only=None, key=None):
        keep = None if only is None else {id(p) for p in only}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None and (keep is None or id(p) in keep):
                    p.add_(p.grad, alpha=-group["lr"])


def _pipelined_case(rank, world):
    """reduce_and_step == reduce() then step(): gradients are slices of one flat buffer (+ one that lives elsewhere, + a
    parameter without gradient), exchanged in 3 pieces with the update of piece i issued while piece i+1 is in flight."""
    from hybrid_ctunet_b200.dp import GradientAllReduce
    shapes = [(3, 5), (7,), (2, 2), (4, 4), (9,), (6,)]

    def build():
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(*s)) for s in shapes] + [torch.nn.Parameter(torch.randn(3))]
        flat = torch.full((64,), float("nan"))
        off = 0
        for i, p in enumerate(params[:5]):
            n = p.numel()
            v = flat[off:off + n].view(p.shape)
            v.copy_(torch.arange(n, dtype=torch.float32).reshape(p.shape) * (rank + 1) + i)
            p.grad = v
            off += -(-n // 4) * 4
        params[5].grad = torch.full((6,), float(rank + 1))      # a gradient outside the flat buffer
        return params                                            # params[6]: no gradient at all
    a, b = build(), build()
    GradientAllReduce(a).reduce_and_step(_SubsetSGD(a, lr=0.5), chunks=3)
    ar = GradientAllReduce(b)
    ar.reduce()
    _SubsetSGD(b, lr=0.5).step()
    return [p.detach().clone() for p in a], [p.detach().clone() for p in b], [None if p.grad is None else p.grad.clone() for p in a]


def test_pipelined_allreduce_and_step_equals_sequential():
    out = _run(_pipelined_case)
    for rank in (0, 1):
        pa, pb, grads = out[rank]
        for x, y in zip(pa, pb):
            assert torch.equal(x, y)
        assert grads[6] is None and torch.equal(grads[5], torch.full((6,), 1.5))
    for x, y in zip(out[0][0], out[1][0]):
        assert torch.equal(x, y)          # both ranks hold the same parameters after the step
