"""Learnable shifts placed before / after every activation quantizer (reference: modules/qbias.py)."""
import torch
import torch.nn as nn


class LearnableBias(nn.Module):
    """qbias.py:5-13. Inside the fused layers only the parameter is used (the add happens in the kernels)."""

    def __init__(self, out_chn):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(out_chn), requires_grad=True)

    def forward(self, x):
        return x + self.bias.expand_as(x)


class LearnableBias4img(nn.Module):
    """qbias.py:15-23: a per-pixel (H*W) shift shared by the three image channels."""

    def __init__(self, out_chn):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(out_chn), requires_grad=True)

    def forward(self, x):
        return x + self.bias.reshape(x.shape[-1], x.shape[-2]).expand_as(x)
