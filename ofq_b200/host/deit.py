"""timm-free fp32 DeiT host model (the L3 "model zoo" layer of SURVEY.md §1).

This is host code around the hot path: patch embedding, class / distillation tokens, LayerNorm, residuals and
heads stay plain PyTorch; `replace_module_by_qmodule_deit` swaps `Attention`, `Mlp`, `patch_embed.proj` and the
heads for the quantized modules.  Module / parameter names follow the reference's deit_vision_transformer.py and
deit.py so that its checkpoints load unchanged (blocks.N.attn.qkv, blocks.N.mlp.fc1, head_dist, ...).
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layers import LayerNorm


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.Identity()

    def forward(self, x):
        return self.norm(self.proj(x).flatten(2).transpose(1, 2))


class Mlp(nn.Module):
    """deit_vision_transformer.py:53-83."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.in_features, self.hidden_features, self.out_features, self.drop = in_features, hidden_features, out_features, drop
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class Attention(nn.Module):
    """deit_vision_transformer.py:85-130 (the qqkkvv diagnostic outputs are not part of the hot path)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0, qqkkvv=False):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.qqkkvv = qqkkvv

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        # unquantized attention (only the KD teacher and un-swapped hosts run it): the library's fused kernel
        x = F.scaled_dot_product_attention(q, k, v, dropout_p=self.attn_drop.p if self.training else 0.0, scale=self.scale)
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x)), None


class Block(nn.Module):
    """deit_vision_transformer.py:132-164."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0,
                 act_layer=nn.GELU, norm_layer=LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self._defer = False       # set by the host model: hand the MLP branch to the next block instead of adding it here

    def forward(self, x, pending=None):
        """pending: the MLP branch of the previous block, not yet added to x (forward_pending): the add is fused into norm1."""
        # ofq_b200 LayerNorm: residual gradient summed in its backward kernel, residual add fused into the forward one
        res, res_add = getattr(self.norm1, "forward_res", None), getattr(self.norm1, "forward_res_add", None)
        if pending is not None:
            x, y = res_add(x, pending) if res_add is not None else (x + pending, None)
            if y is None:
                x, y = res(x) if res is not None else (x, self.norm1(x))
        else:
            x, y = res(x) if res is not None else (x, self.norm1(x))
        a, _ = self.attn(y)
        a = self.drop_path(a)
        res, res_add = getattr(self.norm2, "forward_res", None), getattr(self.norm2, "forward_res_add", None)
        if res_add is not None:
            x, y = res_add(x, a)
        else:
            x = x + a
            x, y = res(x) if res is not None else (x, self.norm2(x))
        m = self.drop_path(self.mlp(y))
        if self._defer:
            return x, m
        return x + m, None


class DistilledVisionTransformer(nn.Module):
    """deit.py:20-67 on top of deit_vision_transformer.py:168-330 (distilled: cls + dist tokens, two heads)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=True, norm_layer=None, act_layer=None):
        super().__init__()
        norm_layer = norm_layer or partial(LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 2
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 2, embed_dim))
        self.pos_drop = nn.Dropout(0.0)
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, qkv_bias, act_layer=act_layer,
                                            norm_layer=norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self.head_dist = nn.Linear(embed_dim, num_classes)
        for t in (self.pos_embed, self.dist_token, self.cls_token):
            nn.init.trunc_normal_(t, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)

    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "dist_token"}

    def forward_features(self, x):
        x = self.patch_embed(x)
        B = x.shape[0]
        x = torch.cat((self.cls_token.expand(B, -1, -1), self.dist_token.expand(B, -1, -1), x), dim=1)
        x = self.pos_drop(x + self.pos_embed)
        # the residual add that closes a block is fused into the first LayerNorm of the next one
        pending = None
        last = len(self.blocks) - 1
        for i, blk in enumerate(self.blocks):
            blk._defer = i < last
            x, pending = blk(x, pending)
        # only the class and distillation tokens reach the heads: LayerNorm is per token, so normalise just those two
        x = self.norm(x[:, :2])
        return x[:, 0], x[:, 1]

    def forward(self, x):
        c, d = self.forward_features(x)
        c, d = self.head(c), self.head_dist(d)
        if self.training:
            return (c, d), None
        return (c + d) / 2, None


def deit_tiny_distilled_patch16_224(**kw):
    return DistilledVisionTransformer(patch_size=16, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True, **kw)


def deit_small_distilled_patch16_224(**kw):
    return DistilledVisionTransformer(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True, **kw)
