"""A plain timm-style distilled DeiT host: torch nn.LayerNorm, un-fused residual adds, the full final norm, and host
classes that are NOT the ones in host/deit.py.

It stands for "somebody else's host model" at the drop-in boundary: `replace_module_by_qmodule_deit` swaps its attention /
MLP / patch-embedding / head modules by interface (modules/utils.py `_qclass_for`), exactly as it would swap the
reference's `src.deit_vision_transformer` classes, and nothing of the repo's host-side glue fusion (fused LayerNorm
kernels, residual add folded into the next norm, two-token final norm) takes part. `bench.py --host-model plain` measures
the quantized modules behind such a host; tests/test_gpu_boundary.py checks it against the reference goldens.
State-dict keys follow deit.py / deit_vision_transformer.py (reference checkpoints load unchanged).
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn


class PlainMlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class PlainAttention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = self.attn_drop(((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1))
        return self.proj_drop(self.proj((attn @ v).transpose(1, 2).reshape(B, N, C)))


class PlainBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, norm_layer=nn.LayerNorm, act_layer=nn.GELU):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = PlainAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = PlainMlp(dim, int(dim * mlp_ratio), act_layer=act_layer)

    def forward(self, x):
        a = self.attn(self.norm1(x))
        x = x + (a[0] if isinstance(a, tuple) else a)          # the quantized attention modules return (x, None)
        return x + self.mlp(self.norm2(x))


class PlainPatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class PlainDistilledViT(nn.Module):
    def __init__(self, img_size=224, patch_size=16, num_classes=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0):
        super().__init__()
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = PlainPatchEmbed(img_size, patch_size, 3, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 2, embed_dim))
        self.blocks = nn.Sequential(*[PlainBlock(embed_dim, num_heads, mlp_ratio, True, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self.head_dist = nn.Linear(embed_dim, num_classes)
        for t in (self.pos_embed, self.dist_token, self.cls_token):
            nn.init.trunc_normal_(t, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "dist_token"}

    def forward(self, x):
        x = self.patch_embed(x)
        B = x.shape[0]
        x = torch.cat((self.cls_token.expand(B, -1, -1), self.dist_token.expand(B, -1, -1), x), dim=1) + self.pos_embed
        x = self.norm(self.blocks(x))
        c, d = self.head(x[:, 0]), self.head_dist(x[:, 1])
        if self.training:
            return (c, d), None
        return (c + d) / 2, None
