"""Drop-in for the reference's networks/vit.py (ViT encoder of TUNet / CTUNet, vit.py:31-139).

Parameter holders keep the reference's registration order and names (`to_patch_embedding.{1,2,3}`,
`pos_embedding`, `transformer.{i}.attn.{norm,to_qkv,to_out.0}`, `transformer.{i}.ff.net.{0,1,4}`); the forward
runs on the sm_100a kernels.  DropPath is constructed but never applied, as in the reference (vit.py:93-96);
dropout > 0 would need RNG parity with torch and is rejected.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._base import KernelModule, Program


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


class DropPath(nn.Module):
    """vit.py:12-29 — kept for constructor parity; the reference never calls it in forward."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return x


def _no_dropout(p: float, what: str):
    if p and p > 0.0:
        raise NotImplementedError(f"{what} > 0 needs torch RNG-stream parity; the BASELINE configs use 0.0")


class FeedForward(KernelModule):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        _no_dropout(dropout, "dropout")
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))

    def _program(self, eng, x):
        t = self._input(x).reshape(-1, x.shape[-1])
        return Program([t], [eng.ffn("", t)], converted=False)

    def forward(self, x):
        """Returns net(x) (without the residual), like vit.py:43-44."""
        return self._call(x)[0].reshape(x.shape) - x


class Attention(KernelModule):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        _no_dropout(dropout, "dropout")
        inner_dim = dim_head * heads
        if heads == 1 and dim_head == dim:
            raise NotImplementedError("project_out=False variant is not used by CTUNet")
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))

    def _program(self, eng, x):
        b, n, d = x.shape
        t = self._input(x).reshape(b * n, d)
        return Program([t], [eng.vit_attention("", t, b, n, self.heads)], converted=False)

    def forward(self, x):
        """Returns the attention branch (without the residual), like vit.py:66-78."""
        return self._call(x)[0].reshape(x.shape) - x


class TransformerBlock(KernelModule):
    def __init__(self, dim, heads, dim_head, mlp_dim, dropout=0.0, drop_path=0.0):
        super().__init__()
        self.attn = Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)
        self.ff = FeedForward(dim, mlp_dim, dropout=dropout)
        self.drop_path = DropPath(dropout) if drop_path > 0.0 else nn.Identity()
        self.heads = heads

    def _program(self, eng, x):
        b, n, d = x.shape
        t = self._input(x).reshape(b * n, d)
        return Program([t], [eng.ffn("ff", eng.vit_attention("attn", t, b, n, self.heads))], converted=False)

    def forward(self, x):
        return self._call(x)[0].reshape(x.shape)


class ViT(KernelModule):
    def __init__(self, image_size, image_patch_size, frames, frame_patch_size, dim, depth, heads, mlp_dim, channels=1,
                 dim_head=64, dropout=0.0, emb_dropout=0.0, drop_path=0.0):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(image_patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, \
            'Image dimensions must be divisible by the patch size.'
        assert frames % frame_patch_size == 0, 'Frames must be divisible by the frame patch size.'
        if channels != 1 or (patch_height, patch_width) != (16, 16) or frame_patch_size not in (8, 16):
            raise NotImplementedError("patchify kernel covers 1 channel, 16x16xpf patches, pf in {8,16}")
        if dim_head != 64:
            raise NotImplementedError("ViT attention kernel is built for dim_head = 64 (vit.py:102 default)")
        _no_dropout(emb_dropout, "emb_dropout")
        num_patches = (image_height // patch_height) * (image_width // patch_width) * (frames // frame_patch_size)
        patch_dim = channels * patch_height * patch_width * frame_patch_size
        self.to_patch_embedding = nn.Sequential(nn.Identity(),  # Rearrange 'b c (h p1) (w p2) (f pf) -> b (h w f) (p1 p2 pf c)'
                                                nn.LayerNorm(patch_dim), nn.Linear(patch_dim, dim), nn.LayerNorm(dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = nn.ModuleList([TransformerBlock(dim, heads, dim_head, mlp_dim, dropout, drop_path)
                                          for _ in range(depth)])
        self.frame_patch_size = frame_patch_size
        self.depth, self.heads, self.dim = depth, heads, dim

    def _program(self, eng, img):
        x, _ = eng.vit("", self._input(img), self.frame_patch_size, self.depth, self.heads)
        return Program([None], [x], converted=False)

    def forward(self, img):
        return self._call(img)[0].view(img.shape[0], -1, self.dim)
