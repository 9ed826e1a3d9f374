"""The drop-in modules under the reference's OWN training / validation harness:

 * fp16 `autocast` + `GradScaler`, exactly the statements of trainer_CTUNet.py:88-109 (`--amp`, the README default);
 * `DistributedDataParallel(find_unused_parameters=True)` as main_CTUNet.py:187-189 wraps the model (two ranks; gloo
   so that both ranks can share the one GPU of the test box — the NCCL path is bench.py's);
 * gradient accumulation / zero_grad(set_to_none=False) (the engine hands out views of a persistent buffer);
 * CUDA-graph inference before and after optimizer steps that do not go through torch's version counter;
 * the Hybrid-CTUNet mask-complementation pipeline (test_CTUNet_final.py:539-552) end to end.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
KW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8)
TKW = dict(in_channels=1, dim_conv_stem=64, out_channels=14, img_size=(96, 96), frames=96, patch_frame=8)


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _loss_func():
    from hybrid_ctunet_b200.losses import DiceCELoss
    return DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)


def test_ctunet_step_under_fp16_autocast_and_gradscaler():
    """trainer_CTUNet.py:88-109 with args.amp: `with autocast(): logits = model(data); loss = ...` then
    `scaler.scale(loss).backward(); scaler.step(optimizer); scaler.update()`.  The kernels compute in bf16 / fp32
    whatever the autocast dtype is; the scaled loss (x 65536) flows through the bf16 activation gradients (fp32 exponent
    range) and is un-scaled on the fp32 parameter gradients: the update must equal the one of the un-scaled step."""
    import scipy.ndimage as ndimage
    from torch.cuda.amp import GradScaler, autocast
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet
    loss_func = _loss_func()
    torch.manual_seed(1)
    data = torch.rand(1, 1, 96, 96, 96, device="cuda")
    target = torch.randint(0, 14, (1, 1, 96, 96, 96), device="cuda").float()

    def step(amp: bool):
        torch.manual_seed(0)
        model = CTUNet(**KW).cuda().train()
        optimizer = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
        scaler = GradScaler() if amp else None
        for param in model.parameters():
            param.grad = None
        with autocast(enabled=amp):
            logits = model(data)
            loss1_1 = loss_func(logits[0][0], target)
            target1 = torch.from_numpy(ndimage.zoom(target.cpu().numpy(), (1, 1, 0.5, 0.5, 1), order=0, prefilter=False)).cuda()
            target2 = torch.from_numpy(ndimage.zoom(target.cpu().numpy(), (1, 1, 0.25, 0.25, 0.5), order=0, prefilter=False)).cuda()
            loss1_2 = loss_func(logits[0][1], target1)
            loss1_3 = loss_func(logits[0][2], target2)
            loss1 = loss1_1 + 0.5 * (loss1_2 + 0.5 * loss1_3)
            loss2 = loss_func(logits[1][0], target) + loss_func(logits[1][1], target)
            loss = loss1 + 0.5 * loss2
        if amp:
            scaler.scale(loss).backward()
            grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
            scaler.step(optimizer)
            scaler.update()
            assert scaler.get_scale() == 65536.0, "GradScaler saw inf/nan gradients and skipped the step"
        else:
            loss.backward()
            grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
            optimizer.step()
        return float(loss.detach()), grads, {n: p.detach().clone() for n, p in model.named_parameters()}

    l_amp, g_amp, p_amp = step(True)
    l_ref, g_ref, p_ref = step(False)
    assert abs(l_amp - l_ref) <= 1e-4 * abs(l_ref)
    assert g_amp.keys() == g_ref.keys() and len(g_ref) == 405
    for n in g_ref:
        assert torch.isfinite(g_amp[n]).all(), n
        # scaled by 2^16 (a power of two: exact in bf16 / fp32) and accumulated in another atomic order
        assert _rel(g_amp[n] / 65536.0, g_ref[n]) < 2e-2, (n, _rel(g_amp[n] / 65536.0, g_ref[n]))
    # AdamW's first step is lr * sign(g) (+ decay): the two runs may only differ where a gradient is ~0
    for n in p_ref:
        assert (p_amp[n] - p_ref[n]).abs().max().item() <= 2.1e-4, n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = H.Up_2Fusion_Block(3, 256, 128, 3, (2, 2, 2), "instance").cuda().train()
        # main_CTUNet.py:187-189
        ddp = DistributedDataParallel(model, device_ids=[0], output_device=0, find_unused_parameters=True)
        res = {}
        for it in range(2):                                   # the second iteration fails if a reduction was left open
            g = torch.Generator(device="cuda").manual_seed(100 * it + rank)
            inp = torch.randn(1, 256, 3, 3, 6, device="cuda", generator=g)
            sc = torch.randn(1, 128, 6, 6, 12, device="cuda", generator=g)
            sv = torch.randn(1, 128, 6, 6, 12, device="cuda", generator=g)
            for p in ddp.parameters():
                p.grad = None
            ddp(inp, sc, sv).square().mean().backward()
            res[it] = {n: (None if p.grad is None else p.grad.detach().cpu()) for n, p in model.named_parameters()}
        torch.save(res, out + f".{rank}")
    finally:
        dist.destroy_process_group()


def test_dropin_under_distributed_data_parallel(tmp_path):
    import torch.multiprocessing as mp
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    out = str(tmp_path / "ddp")
    mp.spawn(_ddp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    # single-process reference: the mean of the two ranks' gradients
    torch.manual_seed(0)
    model = H.Up_2Fusion_Block(3, 256, 128, 3, (2, 2, 2), "instance").cuda().train()
    for it in range(2):
        acc = {}
        for rank in range(2):
            g = torch.Generator(device="cuda").manual_seed(100 * it + rank)
            inp = torch.randn(1, 256, 3, 3, 6, device="cuda", generator=g)
            sc = torch.randn(1, 128, 6, 6, 12, device="cuda", generator=g)
            sv = torch.randn(1, 128, 6, 6, 12, device="cuda", generator=g)
            for p in model.parameters():
                p.grad = None
            model(inp, sc, sv).square().mean().backward()
            for n, p in model.named_parameters():
                if p.grad is not None:
                    acc[n] = acc.get(n, 0) + p.grad.detach().cpu() / 2
        unused = [n for n in r0[it] if r0[it][n] is None]
        # conv3 of the two ResBlocks with equal in/out channels: unreachable in the reference's graph, so DDP leaves
        # their gradient None (and AdamW never decays them) — same here
        assert sorted(unused) == ["up_addconv_block1.conv3.conv.weight", "up_addconv_block2.conv3.conv.weight"]
        for n, g in acc.items():
            assert r0[it][n] is not None and torch.equal(r0[it][n], r1[it][n]), n      # all-reduced: identical on both
            assert _rel(r0[it][n], g) < 1e-3, (it, n, _rel(r0[it][n], g))


def test_gradient_accumulation_over_two_backward_passes():
    """p.grad after two backward passes without zeroing == g1 + g2 (the engine returns views of a persistent buffer;
    AccumulateGrad keeps the first one as .grad without copying)."""
    from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
    torch.manual_seed(0)
    blk = H.ResBlock(3, 128, 64, 3, 1, "instance").cuda().train()
    xs = [torch.randn(1, 128, 8, 12, 16, device="cuda") for _ in range(2)]
    singles = []
    for x in xs:
        for p in blk.parameters():
            p.grad = None
        blk(x).square().mean().backward()
        singles.append({n: p.grad.clone() for n, p in blk.named_parameters()})
    for p in blk.parameters():
        p.grad = None
    for x in xs:
        blk(x).square().mean().backward()
    for n, p in blk.named_parameters():
        want = singles[0][n] + singles[1][n]
        assert _rel(p.grad, want) < 1e-3, (n, _rel(p.grad, want))
    # and with zero_grad(set_to_none=False): the zeroed .grad tensors still alias the engine's buffer
    blk.zero_grad(set_to_none=False)
    blk(xs[0]).square().mean().backward()
    for n, p in blk.named_parameters():
        assert _rel(p.grad, singles[0][n]) < 1e-3, n


def test_graph_inference_stays_correct_across_training_steps():
    """graph-infer, train with the fused optimizer (writes parameters through raw pointers), eval(),
    graph-infer again — must equal eager inference with the updated weights (the relative-position bias tables and the
    Cin = 1 conv weights the graph reads are rebuilt in place, never re-allocated)."""
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import TUNet
    from hybrid_ctunet_b200.optim import AdamW
    torch.manual_seed(0)
    m = TUNet(**TKW).cuda().eval()
    m.enable_cuda_graph(True)
    torch.manual_seed(1)
    x = torch.rand(1, 1, 96, 96, 96, device="cuda")
    y = torch.randint(0, 14, (1, 1, 96, 96, 96), device="cuda").float()
    with torch.no_grad():
        before = [t.clone() for t in m(x)]
    opt = AdamW(m.parameters(), lr=1e-2, weight_decay=1e-5)
    loss_func = _loss_func()
    m.train()
    for _ in range(2):
        for p in m.parameters():
            p.grad = None
        lg = m(x)
        (loss_func(lg[0], y) + loss_func(lg[1], y)).backward()
        opt.step()
    # pile allocations on top of whatever the refresh may have freed
    junk = [torch.full((1 << 20,), float("nan"), device="cuda") for _ in range(64)]
    m.eval()
    with torch.no_grad():
        graphed = [t.clone() for t in m(x)]
        m.enable_cuda_graph(False)
        eager = m(x)
    del junk
    assert _rel(graphed[0], before[0]) > 1e-2, "the optimizer steps did not change the output: nothing was tested"
    for g, e in zip(graphed, eager):
        assert torch.isfinite(g).all()
        assert _rel(g, e) < 1e-3, _rel(g, e)


def test_hybrid_ctunet_pipeline_matches_oracle_ensemble():
    """Config 5 end to end on a small volume: CTUNet head 0 @ overlap 0.5 + TUNet head 0 @ overlap 0.7 +
    mask-complementation ensemble (test_CTUNet_final.py:539-552), against the oracle blend + reference ensemble arithmetic
    applied to the same models' logits."""
    import numpy as np
    from hybrid_ctunet_b200.ensemble import hybrid_ctunet_inference
    from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet, TUNet
    from oracle import sliding_window_oracle as SO
    torch.manual_seed(0)
    ctunet = CTUNet(**KW).cuda().eval()
    torch.manual_seed(7)
    tunet = TUNet(**TKW).cuda().eval()
    torch.manual_seed(2)
    vol = torch.rand(1, 1, 96, 96, 120, device="cuda")
    lab = torch.randint(0, 14, (96, 96, 120), device="cuda").float()
    got = hybrid_ctunet_inference(vol, ctunet, tunet, labels=lab)
    with torch.no_grad():
        p1 = SO.sliding_window_inference(vol, (96, 96, 96), 4, ctunet, overlap=0.5, mode="gaussian", two_heads=True)[0]
        p2 = SO.sliding_window_inference(vol, (96, 96, 96), 4, tunet, overlap=0.7, mode="gaussian", two_heads=False)
    ref = SO.ensemble_reference(p1[0], p2[0], lab)
    assert got["ensemble"].shape == (96, 96, 120) and got["ensemble"].dtype == torch.uint8
    # two runs of a model differ in the last bits (fp64 atomics of the InstanceNorm statistics): masks agree except
    # where two classes tie to ~1e-6
    for k in ("ensemble", "head1", "head2"):
        diff = (got[k].cpu().numpy() != ref[k]).mean()
        assert diff < 2e-3, (k, diff)
    assert np.allclose(got["dice"].cpu().numpy(), ref["dice"], atol=5e-3)


def _sw_worker(rank, world, port, out):
    import torch.distributed as dist
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pred = lambda w: ((torch.sin(3 * w).repeat(1, 5, 1, 1, 1),), (torch.cos(2 * w).repeat(1, 5, 1, 1, 1),))
        res = {}
        for name, shape in (("one_window", (1, 1, 32, 32, 32)), ("three_windows", (1, 1, 32, 32, 64)),
                            ("nine_windows", (1, 1, 32, 32, 160))):
            g = torch.Generator(device="cuda").manual_seed(9)
            vol = torch.rand(*shape, device="cuda", generator=g)
            a0, a1 = sliding_window_inference(vol, (32, 32, 32), 4, pred, overlap=0.5, mode="gaussian",
                                              shard_group=dist.group.WORLD)
            res[name] = (a0.cpu(), a1.cpu())
        torch.save(res, out + f".{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_sliding_window_with_fewer_windows_than_ranks(tmp_path):
    """a rank that owns no window (volume <= roi on 2 ranks) joins the collectives with zero
    accumulators instead of raising while its peers wait in the all-reduce; every rank returns the full result.  Nine windows
    at sw_batch 4 on 2 ranks: chunks of whole calls (8 + 1 windows, sliding_window.shard_window_range)."""
    import torch.multiprocessing as mp
    from hybrid_ctunet_b200.trainer_CTUNet import sliding_window_inference
    out = str(tmp_path / "sw")
    mp.spawn(_sw_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    pred = lambda w: ((torch.sin(3 * w).repeat(1, 5, 1, 1, 1),), (torch.cos(2 * w).repeat(1, 5, 1, 1, 1),))
    for name, shape in (("one_window", (1, 1, 32, 32, 32)), ("three_windows", (1, 1, 32, 32, 64)),
                            ("nine_windows", (1, 1, 32, 32, 160))):
        g = torch.Generator(device="cuda").manual_seed(9)
        vol = torch.rand(*shape, device="cuda", generator=g)
        s0, s1 = sliding_window_inference(vol, (32, 32, 32), 4, pred, overlap=0.5, mode="gaussian")
        for got in (r0[name], r1[name]):
            assert torch.allclose(got[0], s0.cpu(), rtol=1e-6, atol=1e-6)
            assert torch.allclose(got[1], s1.cpu(), rtol=1e-6, atol=1e-6)
