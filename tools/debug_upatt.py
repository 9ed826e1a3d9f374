import sys, torch
sys.path.insert(0, ".")
import torch.nn.functional as F
from hybrid_ctunet_b200.networks import hybrid_CTUNet as H
from hybrid_ctunet_b200 import ops
from oracle import ctunet_oracle as O
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
for Z in (6, 12):
    torch.manual_seed(31)
    m = H.UpAttentionBlock(3, 768, dims=[128, 256, 512, 1024]).cuda().eval()
    x = torch.randn(1, 768, 6, 6, Z).cuda()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        eng = m._engine()
        B, C, X, Y, _ = x.shape
        tok = x.permute(0, 2, 3, 4, 1).reshape(-1, C).contiguous()
        p = "layers.0.0"
        # oracle, stage 0 step by step (grid == windows of 6: Z=6 -> 1 window, Z=12 -> 2)
        w = 6
        t = x.reshape(B, C, X//w, w, Y//w, w, Z//w, w).permute(0, 2, 4, 6, 3, 5, 7, 1)
        a1 = O.multi_axis_attention(sd, p+".1.fn", t) + t
        f1 = O.feed_forward(sd, p+".2.fn", a1) + a1
        xr = f1.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, C, X, Y, Z)
        t = xr.reshape(B, C, w, X//w, w, Y//w, w, Z//w).permute(0, 3, 5, 7, 2, 4, 6, 1)
        a2 = O.multi_axis_attention(sd, p+".5.fn", t) + t
        f2 = O.feed_forward(sd, p+".6.fn", a2) + a2
        xr2 = f2.permute(0, 7, 4, 1, 5, 2, 6, 3).reshape(B, C, X, Y, Z)
        ps = O.pixel_shuffle(sd, p+".8", xr2, (2, 2, 2))
        cl = lambda v: v.permute(0, 2, 3, 4, 1).reshape(-1, v.shape[1])
        a1r = cl(a1.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, C, X, Y, Z))
        # engine
        xs = tok.clone()
        eng.window_attention(p+".1.fn", xs, (B, X, Y, Z), 1); print(Z, "maa1", rel(xs, a1r))
        eng.ffn(p+".2.fn", xs); print(Z, "ffn1", rel(xs, cl(xr)))
        eng.window_attention(p+".5.fn", xs, (B, X, Y, Z), 2)
        a2r = cl(a2.permute(0, 7, 4, 1, 5, 2, 6, 3).reshape(B, C, X, Y, Z)); print(Z, "maa2", rel(xs, a2r))
        xb = torch.empty(xs.shape, dtype=torch.bfloat16, device="cuda")
        eng.ffn(p+".6.fn", xs, out=xb); print(Z, "ffn2", rel(xb.float(), cl(xr2)))
        y = eng.up_gemm(xb.view(B, X, Y, Z, C), eng.w.pixel_shuffle(p+".8.to_out", (2, 2, 2)))
        print(Z, "ps", rel(y.float(), ps.permute(0, 2, 3, 4, 1)), "ps-on-ref-input",
              rel(eng.up_gemm(cl(xr2).to(torch.bfloat16).view(B, X, Y, Z, C).contiguous(), eng.w.pixel_shuffle(p+".8.to_out", (2, 2, 2))).float(), ps.permute(0, 2, 3, 4, 1)))
        ys = m(x); ref = O.up_attention_block(sd, "", x)
        for i in range(1, 5): print(Z, "stage", i, rel(ys[i], ref[i]))
