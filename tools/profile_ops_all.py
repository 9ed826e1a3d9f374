"""Per-shape CUDA-event timing of EVERY ops.* launch in one CTUNet training step (eager, single stream).

Bytes are the tensors each call touches (every distinct tensor argument once: numel x element size), i.e. the
algorithmic traffic of an HBM-bound kernel; GB/s = bytes / event time.  Output: one line per (op, shapes) class,
sorted by total time.  `python tools/profile_ops_all.py [B] [json_out] [--infer]`.
"""
import collections
import json
import sys
sys.path.insert(0, ".")
import torch
from hybrid_ctunet_b200 import ops
from hybrid_ctunet_b200.losses import DiceCELoss, ctunet_loss
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
torch.manual_seed(0)
m = CTUNet(in_channels=1, dim_conv_stem=64, out_channels=14, model_depth=101, img_size=(96, 96), frames=96, patch_frame=8).cuda().train()
lf = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)
x = torch.rand(B, 1, 96, 96, 96, device="cuda")
y = torch.randint(0, 14, (B, 1, 96, 96, 96), device="cuda").float()


INFER = "--infer" in sys.argv   # forward only, eval mode (one sliding-window call of B windows)
if INFER:
    m.eval()


def step():
    if INFER:
        with torch.no_grad():
            m(x)
        return
    for p in m.parameters():
        p.grad = None
    ctunet_loss(m(x), y, lf).backward()


step()
torch.cuda.synchronize()
rec = []
# phase tagging: which part of the network (ViT branch / ResNet encoder / fusion decoder) and which direction an op
# belongs to — the ViT branch and the encoder run as two concurrent lanes under CUDA-graph capture
from hybrid_ctunet_b200.engine import Engine
PHASE = ["dec", "fwd"]


def _phase(name, fn):
    def f(self, *a, **k):
        old = PHASE[0]
        PHASE[0] = name
        try:
            return fn(self, *a, **k)
        finally:
            PHASE[0] = old
    return f


Engine._vit_branch = _phase("vit", Engine._vit_branch)
Engine.resnet = _phase("resnet", Engine.resnet)
_orig_rec = Engine._rec


def _rec(self, fn):
    ph = PHASE[0]

    def g():
        PHASE[0], PHASE[1] = ph, "bwd"
        fn()
    _orig_rec(self, g)


Engine._rec = _rec
SKIP = {"pick_box", "pick_block_n", "pick_wgrad_block_n", "pack_matrix", "check", "lru_cache", "dataclass"}


def tensors_of(args, kwargs):
    seen, out = set(), []
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, ops.PackedWeight):
            a = a.w
        if isinstance(a, torch.Tensor) and a.is_cuda and a.data_ptr() not in seen:
            seen.add(a.data_ptr())
            out.append(a)
    return out


def wrap(name, fn):
    def f(*args, **kwargs):
        ts = tensors_of(args, kwargs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*args, **kwargs)
        e1.record()
        by = sum(t.numel() * t.element_size() for t in ts)
        key = (name,) + tuple("x".join(map(str, t.shape)) + ":" + str(t.dtype).replace("torch.", "") for t in ts)
        if "dims" in kwargs:
            key += (tuple(kwargs["dims"]),)
        rec.append((key, e0, e1, by, (PHASE[0], PHASE[1])))
        return r
    return f


for name in dir(ops):
    fn = getattr(ops, name)
    if callable(fn) and not name.startswith("_") and name not in SKIP and getattr(fn, "__module__", "") == ops.__name__ \
            and not isinstance(fn, type):
        setattr(ops, name, wrap(name, fn))
PHASE[1] = "fwd"
torch.cuda._sleep(int(3e8))   # park the GPU while the host enqueues the step: events then bracket kernel time only
step()
torch.cuda.synchronize()
agg = collections.OrderedDict()
per_op = collections.OrderedDict()
phases = collections.OrderedDict()
by_phase = collections.OrderedDict()
for k, e0, e1, by, ph in rec:
    ms = e0.elapsed_time(e1)
    pa = phases.setdefault(ph, [0, 0.0])
    pa[0] += 1; pa[1] += ms
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += by
    a = by_phase.setdefault(ph, collections.OrderedDict()).setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += by
    p = per_op.setdefault(k[0], [0, 0.0, 0.0])
    p[0] += 1; p[1] += ms; p[2] += by
tot = sum(a[1] for a in agg.values())
print(f"total op time {tot:.2f} ms over {len(rec)} calls")
for ph, (n, ms) in phases.items():
    print(f"  phase {ph[0]:7s} {ph[1]}: {ms:7.2f} ms x{n:4d}")
for k, (n, ms, by) in sorted(per_op.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms x{n:4d}  {by / ms / 1e6:8.1f} GB/s  {k}")
print()
rows = []
for k, (n, ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:150]:
    print(f"{ms:8.3f} ms x{n:3d} avg {1e3 * ms / n:7.1f} us {by / ms / 1e6:8.1f} GB/s  {k}")
    rows.append({"op": k[0], "args": [str(v) for v in k[1:]], "calls": n, "ms": ms, "gbs": by / ms / 1e6})
for ph, d in by_phase.items():
    print(f"\n== phase {ph[0]} {ph[1]}: top classes")
    for k, (n, ms, by) in sorted(d.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"{ms:8.3f} ms x{n:3d} avg {1e3 * ms / n:7.1f} us {by / ms / 1e6:8.1f} GB/s  {k}")
if len(sys.argv) > 2 and sys.argv[2].endswith(".json"):
    json.dump({"B": B, "total_ms": tot, "classes": rows}, open(sys.argv[2], "w"), indent=1)
