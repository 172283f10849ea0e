"""Quantized linear layers (reference: src/quantization/modules/qlinear.py). Same class names, constructor
keywords, forward signatures, sub-module and parameter names; the arithmetic runs in the sm_100a kernels."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...host.deit import Mlp
from ...ops import ACT_GELU, ACT_NONE
from ..functional import F16 as F16_BWD
from ..functional import MlpLink, PatchEmbedFn, QLinearFn
from ..quantizer.lsq import _eff_scale
from ..quantizer.lsq import (LsqQuantizer, LsqQuantizer4Conv2d, LsqQuantizer4head_input, LsqQuantizer4img,
                             LsqQuantizerWeight)
from ..quantizer.statsq import StatsQuantizer
from .qbias import LearnableBias, LearnableBias4img


class LSQ_input(nn.Module):
    """qlinear.py:12-26: move_b4 -> LsqQuantizer -> move_aft as a stand-alone module. Inside
    QAttention_qkreparam only its parameters are used (the codes never leave the fused kernels)."""

    def __init__(self, bit=2, all_positive=False, learnable=True, learanbaleBiasdim=192):
        super().__init__()
        self.input_bits = bit
        self.all_positive = all_positive
        self.learnable = learnable
        self.input_quant_fn = LsqQuantizer(bit=bit, all_positive=all_positive, learnable=learnable)
        self.move_b4 = LearnableBias(learanbaleBiasdim)
        self.move_aft = LearnableBias(learanbaleBiasdim)

    def forward(self, input):
        return self.move_aft(self.input_quant_fn(self.move_b4(input)))


class QLinear(nn.Linear):
    """qlinear.py:28-87."""

    def __init__(self, *kargs, m: nn.Linear, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 symmetric=True, weight_channelwise=True, input_channelwise=True, weight_quant_method="statsq",
                 input_quant_method="lsq", pretrained_initialized=False, **kwargs):
        super().__init__(m.in_features, m.out_features, bias=True)
        self.weight_bits = weight_bits
        self.input_bits = input_bits
        self.aq_learnable = aq_learnable
        self.wq_learnable = wq_learnable
        self.symmetric = symmetric
        self.weight_channelwise = weight_channelwise
        self.input_channelwise = input_channelwise
        self.weight_quant_method = weight_quant_method
        self.input_quant_method = input_quant_method
        self.input_quant_fn = LsqQuantizer(bit=input_bits, all_positive=(symmetric == False), learnable=aq_learnable)  # noqa: E712
        self.pretrained_initialized = pretrained_initialized
        if pretrained_initialized != False:  # noqa: E712
            self.weight = nn.Parameter(m.weight.detach())
            if m.bias is not None:
                self.bias = nn.Parameter(m.bias.detach())
        if weight_quant_method == "statsq":
            self.statsq_fn = StatsQuantizer(num_bits=self.weight_bits, clip_learnable=wq_learnable).to(m.weight.device)
        else:
            raise ValueError("Unknown quant_method")
        self.move_b4 = LearnableBias(self.weight.shape[1])
        self.move_aft = LearnableBias(self.weight.shape[1])

    def forward(self, input, act: int = ACT_NONE, link=None, role: int = 0):
        """act / link / role are used by QMLP only (GELU fused into this layer's input quantizer, see QLinearFn);
        the reference signature forward(input) is unchanged."""
        if self.weight_quant_method != "statsq":
            raise ValueError("Unknown quant_method")
        q = self.input_quant_fn
        if not q.initialized_alpha:
            seen = F.gelu(input.detach()) if act == ACT_GELU else input.detach()
            q.init_from(seen + self.move_b4.bias)
        return QLinearFn.apply(input, self.weight, self.bias, self.move_b4.bias, self.move_aft.bias, q.s,
                               self.weight_bits, self.input_bits, not self.symmetric, act, link, role)

    def extra_repr(self):
        return (f"act_bit={self.input_bits}, weight_bit={self.weight_bits}, act_all_positive={not self.symmetric}, "
                f"wq_learnable={self.wq_learnable}, aq_learnable={self.aq_learnable}, "
                f"weight_quant_method={self.weight_quant_method}, activation_quant_method={self.input_quant_method}, "
                f"pretrained_initialized = {self.pretrained_initialized}")


class QMLP(Mlp):
    """qlinear.py:89-136: fc1 (signed input) -> GELU -> fc2 (unsigned input)."""

    def __init__(self, *kargs, m: Mlp, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 weight_channelwise=True, input_channelwise=True, weight_quant_method="statsq", input_quant_method="lsq",
                 act_layer=nn.GELU, pretrained_initialized=False, **kwargs):
        # any Mlp-like host module (fc1 / fc2 Linears): the repo's, the reference's (deit_vision_transformer.py:53-83), timm's
        drop = getattr(m, "drop", None)
        drop = drop.p if isinstance(drop, nn.Dropout) else (drop if isinstance(drop, (int, float)) else getattr(getattr(m, "drop1", None), "p", 0.0))
        super().__init__(in_features=getattr(m, "in_features", m.fc1.in_features),
                         hidden_features=getattr(m, "hidden_features", None) or m.fc1.out_features,
                         out_features=getattr(m, "out_features", None) or m.fc2.out_features, drop=drop)
        common = dict(weight_bits=weight_bits, input_bits=input_bits, aq_learnable=aq_learnable, wq_learnable=wq_learnable,
                      weight_channelwise=weight_channelwise, input_channelwise=input_channelwise,
                      weight_quant_method=weight_quant_method, input_quant_method=input_quant_method,
                      pretrained_initialized=pretrained_initialized)
        self.fc1 = QLinear(m=m.fc1, symmetric=True, **common)
        self.act_layer = act_layer
        if act_layer == "rprelu":
            raise NotImplementedError("rprelu activations are not used by any OFQ recipe")
        self.act = act_layer() if act_layer != "None" else nn.Identity()
        self.fc2 = QLinear(m=m.fc2, symmetric=False, **common)

    def forward(self, x):
        return qmlp_forward(self, x)


def qmlp_forward(mlp, x):
    """fc2(drop(GELU(fc1(x)))) (qlinear.py:123-136). With nn.GELU() and no active dropout in between, the activation is
    evaluated inside fc2's input-quantizer kernel (forward) and inside its LSQ backward kernel (GELU'), so the GELU output
    and its gradient never travel through HBM."""
    fused = (type(mlp.act) is nn.GELU and getattr(mlp.act, "approximate", "none") == "none" and x.is_cuda
             and (not mlp.training or getattr(mlp.drop1, "p", 0.0) == 0.0) and type(mlp.fc2) is QLinear)
    if not fused:
        x = mlp.drop1(mlp.act(mlp.fc1(x)))
        return mlp.drop2(mlp.fc2(x))
    link = MlpLink(fuse=True)      # h feeds fc2 only: its gradient travels as fc1's fp16 GEMM operand, not as an fp32 tensor
    h = mlp.fc1(x, ACT_NONE, link, 1)
    return mlp.drop2(mlp.fc2(h, ACT_GELU, link, 2))


class LSQ_QConv2d(nn.Conv2d):
    """qlinear.py:138-191: the 8-bit patch-embedding convolution (torch-composed)."""

    def __init__(self, *kargs, m: nn.Conv2d, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 symmetric=True, weight_channelwise=True, input_channelwise=True, weight_quant_method="lsq",
                 input_quant_method="lsq", pretrained_initialized=False, **kwargs):
        super().__init__(in_channels=m.in_channels, out_channels=m.out_channels, kernel_size=m.kernel_size,
                         stride=m.stride, padding=m.padding, dilation=m.dilation, groups=m.groups, bias=True)
        self.weight_bits, self.input_bits = weight_bits, input_bits
        self.aq_learnable, self.wq_learnable, self.symmetric = aq_learnable, wq_learnable, symmetric
        self.input_quant_fn = LsqQuantizer4img(bit=input_bits, all_positive=(symmetric == False), learnable=aq_learnable)  # noqa: E712
        self.pretrained_initialized = pretrained_initialized
        if pretrained_initialized != False:  # noqa: E712
            self.weight = nn.Parameter(m.weight.detach())
            if m.bias is not None:
                self.bias = nn.Parameter(m.bias.detach())
        self.lsqw_fn = LsqQuantizer4Conv2d(bit=self.weight_bits, learnable=aq_learnable).to(m.weight.device)
        self.move_b4 = LearnableBias4img(224 * 224)
        self.move_aft = LearnableBias4img(224 * 224)

    def forward(self, input):
        weight = self.lsqw_fn(self.weight)
        kh, kw = self.kernel_size
        q = self.input_quant_fn
        patchify = (tuple(self.stride) == (kh, kw) and tuple(self.padding) == (0, 0) and tuple(self.dilation) == (1, 1)
                    and self.groups == 1 and input.shape[-2] % kh == 0 and input.shape[-1] % kw == 0)
        if (patchify and F16_BWD and q.fused_ready(input, self.move_b4, self.move_aft) and (kh * kw) % 16 == 0
                and self.lsqw_fn.initialized_alpha and self.bias is not None):
            # integer codes of image and weight straight into int8 tensor-core GEMMs (functional.PatchEmbedFn)
            w = self.lsqw_fn
            se_w = _eff_scale(w.s.detach(), 1.0 / ((w.thd_pos * weight[0].numel()) ** 0.5))
            return PatchEmbedFn.apply(input, self.move_b4.bias, self.move_aft.bias, q.s, weight, se_w, self.bias,
                                      -(2 ** (q.bit - 1)), 2 ** (q.bit - 1) - 1)
        input = self.input_quant_fn.forward_with_shifts(input, self.move_b4, self.move_aft)
        if (tuple(self.stride) == (kh, kw) and tuple(self.padding) == (0, 0) and tuple(self.dilation) == (1, 1)
                and self.groups == 1 and input.shape[-2] % kh == 0 and input.shape[-1] % kw == 0):
            # patchify convolution == one GEMM over unfolded patches; done as a true-fp32 matmul because cuDNN
            # would pick TF32 kernels (forward and backward) and the 8-bit operands would no longer be exact
            B, Cin, Hh, Ww = input.shape
            cols = input.view(B, Cin, Hh // kh, kh, Ww // kw, kw).permute(0, 2, 4, 1, 3, 5).reshape(-1, Cin * kh * kw)
            out = torch.addmm(self.bias, cols, weight.view(weight.shape[0], -1).t())
            return out.view(B, Hh // kh, Ww // kw, -1).permute(0, 3, 1, 2)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return F.conv2d(input, weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class LSQ_QLinear4head(nn.Linear):
    """qlinear.py:193-252: the 8-bit classifier heads: int8 tensor-core path once the step
    sizes exist, the op-for-op torch composition for the initialising forward and for shapes / options outside that path."""

    def __init__(self, *kargs, m: nn.Linear, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 symmetric=True, weight_channelwise=True, input_channelwise=True, weight_quant_method="statsq",
                 input_quant_method="lsq", pretrained_initialized=False, **kwargs):
        super().__init__(m.in_features, m.out_features, bias=True)
        self.weight_bits, self.input_bits = weight_bits, input_bits
        self.aq_learnable, self.wq_learnable, self.symmetric = aq_learnable, wq_learnable, symmetric
        self.weight_quant_method = weight_quant_method
        self.input_quant_fn = LsqQuantizer4head_input(bit=input_bits, all_positive=(symmetric == False), learnable=aq_learnable)  # noqa: E712
        self.pretrained_initialized = pretrained_initialized
        if pretrained_initialized != False:  # noqa: E712
            self.weight = nn.Parameter(m.weight.detach())
            if m.bias is not None:
                self.bias = nn.Parameter(m.bias.detach())
        if weight_quant_method == "lsq":
            self.lsqw_fn = LsqQuantizerWeight(bit=self.weight_bits, per_channel=weight_channelwise, learnable=wq_learnable).to(m.weight.device)
        else:
            raise ValueError("Unknown quant_method")
        self.move_b4 = LearnableBias(self.weight.shape[1])
        self.move_aft = LearnableBias(self.weight.shape[1])

    def _native(self, input) -> bool:
        """The int8 tensor-core path (functional.HeadLinearFn): CUDA fp32, signed 8-bit codes, both step sizes already created
        (the first forward initialises them from the data, lsq.py:72-101 / 476-486) and TMA-legal pitches."""
        return (input.is_cuda and input.dtype == torch.float32 and self.symmetric and self.weight_bits == 8 and self.input_bits == 8
                and self.input_quant_fn.initialized_alpha and self.lsqw_fn.initialized_alpha
                and self.weight.shape[0] % 8 == 0 and self.weight.shape[1] % 16 == 0 and input.shape[-1] == self.weight.shape[1])

    def forward(self, input):
        if self._native(input):
            from ..functional import HeadLinearFn
            return HeadLinearFn.apply(input, self.weight, self.bias, self.move_b4.bias, self.move_aft.bias,
                                      self.input_quant_fn.s, self.lsqw_fn.s, self.weight_bits)
        weight = self.lsqw_fn(self.weight)
        input = self.move_aft(self.input_quant_fn(self.move_b4(input)))
        out = F.linear(input, weight)
        return out + self.bias.view(1, -1).expand_as(out)
