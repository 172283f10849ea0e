"""timm/torchvision-free fp32 Swin host model. Module names follow torchvision's / the reference's
src/swin.py (`features.N.M.attn.qkv`, `features.N.reduction`, `head`) so that reference checkpoints load unchanged.
Host code around the hot path: patch embedding, LayerNorm, residuals, patch merging, pooling and the head are PyTorch;
`replace_module_by_qmodule_swin` swaps `ShiftedWindowAttention`, `MLP`, `reduction`, `features.0.0` and `head`."""
from __future__ import annotations

from functools import partial
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layers import LayerNorm


class Permute(nn.Module):
    def __init__(self, dims):
        super().__init__()
        self.dims = dims

    def forward(self, x):
        return x.permute(*self.dims)


class MLP(nn.Sequential):
    """torchvision.ops.misc.MLP layout: Linear, GELU, Dropout, Linear, Dropout."""

    def __init__(self, in_channels: int, hidden_channels: List[int], dropout: float = 0.0):
        layers = []
        d = in_channels
        for h in hidden_channels[:-1]:
            layers += [nn.Linear(d, h), nn.GELU(), nn.Dropout(dropout)]
            d = h
        layers += [nn.Linear(d, hidden_channels[-1]), nn.Dropout(dropout)]
        super().__init__(*layers)


def relative_position_index(window_size):
    """src/swin.py:203-213."""
    coords = torch.stack(torch.meshgrid(torch.arange(window_size[0]), torch.arange(window_size[1]), indexing="ij"))
    flat = torch.flatten(coords, 1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window_size[0] - 1
    rel[:, :, 1] += window_size[1] - 1
    rel[:, :, 0] *= 2 * window_size[1] - 1
    return rel.sum(-1).view(-1)


def shift_mask(pad_H, pad_W, window_size, shift_size, device):
    """0 / -100 mask of the cyclically shifted windows (src/swin.py:118-133), [nW, N, N]."""
    m = torch.zeros((pad_H, pad_W), device=device)
    hs = ((0, -window_size[0]), (-window_size[0], -shift_size[0]), (-shift_size[0], None))
    ws = ((0, -window_size[1]), (-window_size[1], -shift_size[1]), (-shift_size[1], None))
    count = 0
    for h in hs:
        for w in ws:
            m[h[0]:h[1], w[0]:w[1]] = count
            count += 1
    m = m.view(pad_H // window_size[0], window_size[0], pad_W // window_size[1], window_size[1])
    m = m.permute(0, 2, 1, 3).reshape(-1, window_size[0] * window_size[1])
    m = m.unsqueeze(1) - m.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


class ShiftedWindowAttention(nn.Module):
    """src/swin.py:176-252 (fp32 host version; returns (x, None) like the reference)."""

    def __init__(self, dim, window_size, shift_size, num_heads, qkv_bias=True, proj_bias=True, attention_dropout=0.0,
                 dropout=0.0, qqkkvv=False):
        super().__init__()
        self.dim, self.window_size, self.shift_size, self.num_heads = dim, list(window_size), list(shift_size), num_heads
        self.attention_dropout, self.dropout, self.qqkkvv = attention_dropout, dropout, qqkkvv
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        self.register_buffer("relative_position_index", relative_position_index(window_size))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def relative_position_bias(self):
        N = self.window_size[0] * self.window_size[1]
        return self.relative_position_bias_table[self.relative_position_index].view(N, N, -1).permute(2, 0, 1).contiguous()

    def windows(self, x):
        """pad -> cyclic shift -> partition. Returns (windows [B*nW, N, C], context for `unwindows`, mask | None, nW)."""
        B, H, W, C = x.shape
        ws = self.window_size
        pad_r, pad_b = (ws[1] - W % ws[1]) % ws[1], (ws[0] - H % ws[0]) % ws[0]
        x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
        _, pH, pW, _ = x.shape
        shift = list(self.shift_size)
        if ws[0] >= pH:
            shift[0] = 0
        if ws[1] >= pW:
            shift[1] = 0
        if sum(shift) > 0:
            x = torch.roll(x, shifts=(-shift[0], -shift[1]), dims=(1, 2))
        nW = (pH // ws[0]) * (pW // ws[1])
        x = x.view(B, pH // ws[0], ws[0], pW // ws[1], ws[1], C).permute(0, 1, 3, 2, 4, 5).reshape(B * nW, ws[0] * ws[1], C)
        mask = shift_mask(pH, pW, ws, shift, x.device) if sum(shift) > 0 else None
        return x, (B, H, W, C, pH, pW, shift), mask, nW

    def unwindows(self, x, ctx):
        B, H, W, C, pH, pW, shift = ctx
        ws = self.window_size
        x = x.view(B, pH // ws[0], pW // ws[1], ws[0], ws[1], C).permute(0, 1, 3, 2, 4, 5).reshape(B, pH, pW, C)
        if sum(shift) > 0:
            x = torch.roll(x, shifts=(shift[0], shift[1]), dims=(1, 2))
        return x[:, :H, :W, :].contiguous()

    def forward(self, x):
        xw, ctx, mask, nW = self.windows(x)
        Bw, N, C = xw.shape
        H = self.num_heads
        qkv = self.qkv(xw).reshape(Bw, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q * (C // H) ** -0.5) @ k.transpose(-2, -1) + self.relative_position_bias().unsqueeze(0)
        if mask is not None:
            attn = (attn.view(Bw // nW, nW, H, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, H, N, N)
        out = (F.softmax(attn, dim=-1) @ v).transpose(1, 2).reshape(Bw, N, C)
        return self.unwindows(self.proj(out), ctx), None


class PatchMerging(nn.Module):
    """src/swin.py:26-59."""

    def __init__(self, dim, norm_layer=LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):
        H, W, _ = x.shape[-3:]
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
        x = torch.cat([x[..., 0::2, 0::2, :], x[..., 1::2, 0::2, :], x[..., 0::2, 1::2, :], x[..., 1::2, 1::2, :]], -1)
        return self.reduction(self.norm(x))


class SwinTransformerBlock(nn.Module):
    """src/swin.py:255-322."""

    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio=4.0, norm_layer=LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = ShiftedWindowAttention(dim, window_size, shift_size, num_heads)
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(dim, [int(dim * mlp_ratio), dim])
        for m in self.mlp.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.normal_(m.bias, std=1e-6)

    def forward(self, x):
        a, _ = self.attn(self.norm1(x))
        x = x + a
        return x + self.mlp(self.norm2(x))


class SwinTransformer(nn.Module):
    """src/swin.py:325-448."""

    def __init__(self, patch_size=(4, 4), embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=(7, 7),
                 mlp_ratio=4.0, num_classes=1000):
        super().__init__()
        norm_layer = partial(LayerNorm, eps=1e-5)
        layers: List[nn.Module] = [nn.Sequential(
            nn.Conv2d(3, embed_dim, kernel_size=tuple(patch_size), stride=tuple(patch_size)), Permute([0, 2, 3, 1]), norm_layer(embed_dim))]
        for i, depth in enumerate(depths):
            dim = embed_dim * 2 ** i
            layers.append(nn.Sequential(*[
                SwinTransformerBlock(dim, num_heads[i], list(window_size), [0 if j % 2 == 0 else w // 2 for w in window_size],
                                     mlp_ratio, norm_layer) for j in range(depth)]))
            if i < len(depths) - 1:
                layers.append(PatchMerging(dim, norm_layer))
        self.features = nn.Sequential(*layers)
        nf = embed_dim * 2 ** (len(depths) - 1)
        self.norm = norm_layer(nf)
        self.permute = Permute([0, 3, 1, 2])
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.flatten = nn.Flatten(1)
        self.head = nn.Linear(nf, num_classes)
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear) and ".mlp." not in name:
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.features(x)
        x = self.flatten(self.avgpool(self.permute(self.norm(x))))
        return self.head(x), None


def swin_t(**kw):
    return SwinTransformer(patch_size=(4, 4), embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=(7, 7), **kw)
