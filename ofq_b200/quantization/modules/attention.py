"""Quantized DeiT attention (reference: src/quantization/modules/attention.py). Same class names, constructor
keywords, forward signatures `forward(x) -> (x, None)`, sub-module and parameter names."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...host.deit import Attention as deit_attention
from ...ops import ACT_NONE
from ..functional import MlpLink, QAttnCoreFn, QKRAttnCoreFn
from ..quantizer.lsq import LsqQuantizer, LsqQuantizer4v
from ..quantizer.statsq import StatsQuantizer, StatsQuantizer_specific_4_qkreparam_cga
from .qbias import LearnableBias
from .qlinear import LSQ_input, QLinear


def _check_attention_like(m):
    """The registry hands over the host model's attention module (modules/utils.py:21-44). Any module with the timm / DeiT
    attention interface is accepted - the repo's own host class, the reference's src.deit_vision_transformer.Attention, or
    timm's - not one particular class (the reference asserts `type(m) == deit_attention`, attention.py:17, because its
    registry is keyed on its own class)."""
    for attr in ("qkv", "proj", "num_heads", "attn_drop", "proj_drop"):
        if not hasattr(m, attr):
            raise TypeError(f"{type(m).__name__} does not look like a DeiT attention module: no `{attr}`")
    return m


def _qlinear_kwargs(weight_bits, input_bits, weight_channelwise, input_channelwise, weight_quant_method,
                    input_quant_method, aq_learnable, wq_learnable, pretrained_initialized):
    return dict(weight_bits=weight_bits, input_bits=input_bits, weight_channelwise=weight_channelwise,
                input_channelwise=input_channelwise, weight_quant_method=weight_quant_method,
                input_quant_method=input_quant_method, aq_learnable=aq_learnable, wq_learnable=wq_learnable,
                symmetric=True, pretrained_initialized=pretrained_initialized)


class QAttention(deit_attention):
    """attention.py:12-105: quantized q, k (per token), v (per channel) and attention probabilities."""

    def __init__(self, m: deit_attention, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 weight_channelwise=True, input_channelwise=True, weight_quant_method="statsq", input_quant_method="lsq",
                 pretrained_initialized=False, **kwargs):
        _check_attention_like(m)
        super().__init__(dim=m.qkv.in_features, num_heads=m.num_heads, attn_drop=m.attn_drop.p, proj_drop=m.proj_drop.p,
                         qqkkvv=getattr(m, "qqkkvv", False))
        self.weight_bits = weight_bits
        self.input_bits = input_bits
        self.input_channelwise = input_channelwise
        kw = _qlinear_kwargs(weight_bits, input_bits, weight_channelwise, input_channelwise, weight_quant_method,
                             input_quant_method, aq_learnable, wq_learnable, pretrained_initialized)
        # as in the reference (attention.py:29,41) the wrapped layers are this module's own freshly constructed
        # self.qkv / self.proj, not m.qkv / m.proj: pretrained values arrive through the checkpoint
        self.qkv = QLinear(m=self.qkv, **kw)
        self.proj = QLinear(m=self.proj, **kw)
        if m.attn_drop.p > 0:
            raise NotImplementedError("attention-probability dropout > 0 is not used by any OFQ recipe")
        dim = m.qkv.in_features
        self.quan_a_q_fn = LsqQuantizer(bit=input_bits, all_positive=False, per_channel=True, learnable=aq_learnable)
        self.quan_a_k_fn = LsqQuantizer(bit=input_bits, all_positive=False, per_channel=True, learnable=aq_learnable)
        self.quan_a_v_fn = LsqQuantizer4v(bit=input_bits, all_positive=False, per_channel=True, learnable=aq_learnable)
        self.move_qkv_b4 = LearnableBias(dim * 3)
        self.move_q_aft = LearnableBias(dim)
        self.move_k_aft = LearnableBias(dim)
        self.move_v_aft = LearnableBias(dim)
        self.quan_a_softmax_fn = LsqQuantizer(bit=input_bits, all_positive=True, per_channel=True, learnable=aq_learnable)

    def _scales_ready(self):
        return all(q.initialized_alpha for q in (self.quan_a_q_fn, self.quan_a_k_fn, self.quan_a_v_fn, self.quan_a_softmax_fn))

    @torch.no_grad()
    def _init_scales(self, qkv, attn_bias=None):
        """Data-dependent creation of the step sizes on the first forward (lsq.py:544-569): a plain staged
        evaluation of attention.py:70-99 that stops at each un-initialised quantizer."""
        B, N, C3 = qkv.shape
        C, H = C3 // 3, self.num_heads
        t = (qkv + self.move_qkv_b4.bias).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
        q, k, v = t[0], t[1], t[2]
        q = self.quan_a_q_fn(q)
        k = self.quan_a_k_fn(k)
        v = self.quan_a_v_fn(v.permute(0, 2, 1, 3).reshape(B, N, C))
        q = (q.permute(0, 2, 1, 3).reshape(B, N, C) + self.move_q_aft.bias).reshape(B, N, H, C // H).permute(0, 2, 1, 3)
        k = (k.permute(0, 2, 1, 3).reshape(B, N, C) + self.move_k_aft.bias).reshape(B, N, H, C // H).permute(0, 2, 1, 3)
        attn = (q @ k.transpose(-2, -1)) * self.scale
        if attn_bias is not None:
            attn = attn + attn_bias
        self.quan_a_softmax_fn(F.softmax(attn, dim=-1))

    def _core(self, qkv, attn_bias=None, attn_mask=None, nW=0, link=None):
        if not self._scales_ready():
            full_bias = attn_bias
            if attn_mask is not None:
                B = qkv.shape[0]
                m = attn_mask[torch.arange(B, device=qkv.device) % nW].unsqueeze(1)
                full_bias = m if attn_bias is None else attn_bias.unsqueeze(0) + m
            self._init_scales(qkv.detach(), full_bias)
        return QAttnCoreFn.apply(qkv, self.move_qkv_b4.bias, self.quan_a_q_fn.s, self.quan_a_k_fn.s, self.quan_a_v_fn.s,
                                 self.move_q_aft.bias, self.move_k_aft.bias, self.move_v_aft.bias,
                                 self.quan_a_softmax_fn.s, attn_bias, attn_mask, self.num_heads, self.input_bits, nW, link)

    def forward(self, x):
        qkv = self.qkv(x)
        link = MlpLink()        # proj's backward hands the fp16 range scale of d(core output) to the core's backward
        x = self._core(qkv, link=link)
        x = self.proj(x, ACT_NONE, link, 2)
        x = self.proj_drop(x)
        return x, None


class QAttention_qkreparam(deit_attention):
    """attention.py:107-222: query-key reparameterisation — one StatsQ on W_q^T W_k, x quantized once and shared
    by the V, W_qk x and score GEMMs."""

    _qk_quant_cls = StatsQuantizer

    def __init__(self, m: deit_attention, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 weight_channelwise=True, input_channelwise=True, weight_quant_method="statsq", input_quant_method="lsq",
                 pretrained_initialized=False, **kwargs):
        _check_attention_like(m)
        super().__init__(dim=m.qkv.in_features, num_heads=m.num_heads, attn_drop=m.attn_drop.p, proj_drop=m.proj_drop.p,
                         qqkkvv=getattr(m, "qqkkvv", False))
        dim = m.qkv.in_features
        self.weight_bits = weight_bits
        self.input_bits = input_bits
        self.input_channelwise = input_channelwise
        self.quant_x_4_qkv = LSQ_input(bit=input_bits, all_positive=False, learnable=aq_learnable, learanbaleBiasdim=dim)
        self.q = nn.Linear(dim, dim, bias=False)
        self.k = nn.Linear(dim, dim, bias=False)
        self.v = nn.Linear(dim, dim)
        if pretrained_initialized:
            with torch.no_grad():
                d = int(m.qkv.weight.shape[0] / 3)
                w, b = m.qkv.weight.detach(), m.qkv.bias.detach()
                self.q.weight.copy_(w[:d])
                self.k.weight.copy_(w[d:2 * d])
                self.v.weight.copy_(w[2 * d:])
                self.v.bias.copy_(b[2 * d:])
        self.qk_quant = self._make_qk_quant(wq_learnable, kwargs)
        self.v_quant = StatsQuantizer(num_bits=self.weight_bits, clip_learnable=wq_learnable)
        kw = _qlinear_kwargs(weight_bits, input_bits, weight_channelwise, input_channelwise, weight_quant_method,
                             input_quant_method, aq_learnable, wq_learnable, pretrained_initialized)
        self.proj = QLinear(m=self.proj, **kw)        # attention.py:142: wraps its own fresh proj
        if m.attn_drop.p > 0:
            raise NotImplementedError("attention-probability dropout > 0 is not used by any OFQ recipe")
        self.quan_a_qkx_fn = LsqQuantizer(bit=input_bits, all_positive=False, per_channel=True, learnable=aq_learnable)
        self.quan_a_v_fn = LsqQuantizer4v(bit=input_bits, all_positive=False, per_channel=True, learnable=aq_learnable)
        self.move_qkx_b4 = LearnableBias(self.num_heads * dim)
        self.move_qkx_aft = LearnableBias(self.num_heads * dim)
        self.move_v_b4 = LearnableBias(dim)
        self.move_v_aft = LearnableBias(dim)
        self.quan_a_softmax_fn = LsqQuantizer(bit=input_bits, all_positive=True, per_channel=True, learnable=aq_learnable)
        del self.qkv

    def _make_qk_quant(self, wq_learnable, kwargs):
        return StatsQuantizer(num_bits=self.weight_bits, clip_learnable=wq_learnable)

    def _scales_ready(self):
        return all(q.initialized_alpha for q in (self.quant_x_4_qkv.input_quant_fn, self.quan_a_qkx_fn,
                                                 self.quan_a_v_fn, self.quan_a_softmax_fn))

    @torch.no_grad()
    def _init_scales(self, x, attn_bias=None):
        """First-forward creation of the step sizes: a staged evaluation of attention.py:174-215."""
        B, N, C = x.shape
        H = self.num_heads
        xq = self.quant_x_4_qkv(x)
        v = F.linear(xq, self.v_quant(self.v.weight)) + self.v.bias + self.move_v_b4.bias
        self.quan_a_v_fn(v)
        wqk = self.qk_quant(ops.wqk_compose(self.q.weight.contiguous(), self.k.weight.contiguous(), H))   # [H*C, C]
        qkx = F.linear(xq, wqk) + self.move_qkx_b4.bias                        # [B, N, H*C]
        qkx = self.quan_a_qkx_fn(qkx.reshape(B, N * H, C)).reshape(B, N, H * C) + self.move_qkx_aft.bias
        attn = torch.einsum("bnc,bdhc->bhnd", xq, qkx.reshape(B, N, H, C)) * self.scale
        if attn_bias is not None:
            attn = attn + attn_bias
        self.quan_a_softmax_fn(F.softmax(attn, dim=-1))

    def _core(self, x, attn_bias=None, attn_mask=None, nW=0, link=None):
        if not self._scales_ready():
            full_bias = attn_bias
            if attn_mask is not None:
                B = x.shape[0]
                m = attn_mask[torch.arange(B, device=x.device) % nW].unsqueeze(1)
                full_bias = m if attn_bias is None else attn_bias.unsqueeze(0) + m
            self._init_scales(x.detach(), full_bias)
        qx = self.quant_x_4_qkv
        return QKRAttnCoreFn.apply(x, self.q.weight, self.k.weight, self.v.weight, self.v.bias,
                                   qx.move_b4.bias, qx.move_aft.bias, qx.input_quant_fn.s,
                                   self.move_v_b4.bias, self.move_v_aft.bias, self.quan_a_v_fn.s,
                                   self.move_qkx_b4.bias, self.move_qkx_aft.bias, self.quan_a_qkx_fn.s,
                                   self.quan_a_softmax_fn.s, attn_bias, attn_mask, self.num_heads, self.weight_bits,
                                   self.input_bits, nW, link)

    def forward(self, x):
        link = MlpLink()        # proj's backward hands the fp16 range scale of d(core output) to the core's backward
        x = self._core(x, link=link)
        x = self.proj(x, ACT_NONE, link, 2)
        x = self.proj_drop(x)
        return x, None


class QAttention_qkreparam_4_cga(QAttention_qkreparam):
    """attention.py:224-339: identical in value and gradient to QAttention_qkreparam (the special StatsQ variant
    is a functional no-op); kept as a class of its own for the registry and checkpoints."""

    def __init__(self, m: deit_attention, clip_val=2.5, boundaryRange=0.005, **kwargs):
        self._boundaryRange = boundaryRange
        super().__init__(m, **kwargs)

    def _make_qk_quant(self, wq_learnable, kwargs):
        return StatsQuantizer_specific_4_qkreparam_cga(num_bits=self.weight_bits, clip_learnable=wq_learnable,
                                                       boundaryRange=self._boundaryRange)
