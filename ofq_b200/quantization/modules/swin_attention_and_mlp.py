"""Quantized Swin attention / MLP (reference: src/quantization/modules/swin_attention_and_mlp.py). Same class names,
constructor keywords, `forward(x[B,H,W,C]) -> (x, None)`, sub-module and parameter names. The attention cores are the
DeiT ones (same kernels) on 49-token windows, with the relative-position bias and the shifted-window mask added before
the softmax inside `ofq_softmax_quant`; pad / cyclic roll / window partition are index shuffles left to torch."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...host.swin import MLP as swin_MLP
from ...host.swin import ShiftedWindowAttention
from ..quantizer.lsq import LsqQuantizer, LsqQuantizer4v
from ..quantizer.statsq import StatsQuantizer, StatsQuantizer_specific_4_qkreparam_cga
from .attention import QAttention, QAttention_qkreparam, _qlinear_kwargs
from .qbias import LearnableBias
from .qlinear import LSQ_input, QLinear, qmlp_forward


def _check_window_attention_like(m):
    """Any module with the torchvision / reference ShiftedWindowAttention interface is accepted (the reference asserts
    `type(m) == ShiftedWindowAttention`, swin_attention_and_mlp.py:71, because its registry is keyed on its own class)."""
    for attr in ("qkv", "proj", "num_heads", "window_size", "shift_size"):
        if not hasattr(m, attr):
            raise TypeError(f"{type(m).__name__} does not look like a shifted-window attention module: no `{attr}`")
    if not hasattr(m, "dim"):
        m.dim = m.qkv.in_features
    return m


class QMLP_swin(nn.Module):
    """swin_attention_and_mlp.py:24-63: fc1 (signed input) -> GELU -> fc2 (unsigned input)."""

    def __init__(self, *kargs, m: swin_MLP, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 weight_channelwise=True, input_channelwise=True, weight_quant_method="statsq", input_quant_method="lsq",
                 act_layer=nn.GELU, pretrained_initialized=False, **kwargs):
        super().__init__()
        common = dict(weight_bits=weight_bits, input_bits=input_bits, aq_learnable=aq_learnable, wq_learnable=wq_learnable,
                      weight_channelwise=weight_channelwise, input_channelwise=input_channelwise,
                      weight_quant_method=weight_quant_method, input_quant_method=input_quant_method,
                      pretrained_initialized=pretrained_initialized)
        self.fc1 = QLinear(m=m[0], symmetric=True, **common)
        self.act_layer = act_layer
        if act_layer == "rprelu":
            raise NotImplementedError("rprelu activations are not used by any OFQ recipe")
        self.act = act_layer() if act_layer != "None" else nn.Identity()
        self.drop1 = m[2]
        self.fc2 = QLinear(m=m[3], symmetric=False, **common)
        self.drop2 = m[4]

    def forward(self, x):
        return qmlp_forward(self, x)


class _SwinWindows:
    """Window plumbing shared by the Swin attention variants (swin_attention_and_mlp.py:143-170, 231-240)."""

    def _window_forward(self, x, core_input_fn):
        xw, ctx, mask, nW = self.windows(x)
        out = self._core(core_input_fn(xw), self.relative_position_bias(), mask, nW if mask is not None else 0)
        out = self.proj(out)
        return self.unwindows(out, ctx), None


class QAttention_swin(_SwinWindows, ShiftedWindowAttention):
    """swin_attention_and_mlp.py:65-251."""

    _core = QAttention._core
    _scales_ready = QAttention._scales_ready
    _init_scales = QAttention._init_scales

    def __init__(self, m: ShiftedWindowAttention, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 weight_channelwise=True, input_channelwise=True, weight_quant_method="statsq", input_quant_method="lsq",
                 pretrained_initialized=False, **kwargs):
        _check_window_attention_like(m)
        super().__init__(dim=m.dim, window_size=m.window_size, shift_size=m.shift_size, num_heads=m.num_heads, qkv_bias=True,
                         proj_bias=True, attention_dropout=0.0, dropout=0.0, qqkkvv=getattr(m, "qqkkvv", False))
        self.weight_bits, self.input_bits, self.input_channelwise = weight_bits, input_bits, input_channelwise
        self.scale = (m.dim // m.num_heads) ** -0.5
        kw = _qlinear_kwargs(weight_bits, input_bits, weight_channelwise, input_channelwise, weight_quant_method,
                             input_quant_method, aq_learnable, wq_learnable, pretrained_initialized)
        self.qkv = QLinear(m=self.qkv, **kw)       # wraps this module's own fresh layers, as the reference does (:87, :99)
        self.proj = QLinear(m=self.proj, **kw)
        dim = m.qkv.in_features
        self.quan_a_q_fn = LsqQuantizer(bit=input_bits, all_positive=False, per_channel=True, learnable=aq_learnable)
        self.quan_a_k_fn = LsqQuantizer(bit=input_bits, all_positive=False, per_channel=True, learnable=aq_learnable)
        self.quan_a_v_fn = LsqQuantizer4v(bit=input_bits, all_positive=False, per_channel=True, learnable=aq_learnable)
        self.move_qkv_b4 = LearnableBias(dim * 3)
        self.move_q_aft = LearnableBias(dim)
        self.move_k_aft = LearnableBias(dim)
        self.move_v_aft = LearnableBias(dim)
        self.quan_a_softmax_fn = LsqQuantizer(bit=input_bits, all_positive=True, per_channel=True, learnable=aq_learnable)

    def forward(self, x):
        return self._window_forward(x, lambda xw: self.qkv(xw))


class QAttention_swin_qkreparam(_SwinWindows, ShiftedWindowAttention):
    """swin_attention_and_mlp.py:253-461."""

    _core = QAttention_qkreparam._core
    _scales_ready = QAttention_qkreparam._scales_ready
    _init_scales = QAttention_qkreparam._init_scales

    def __init__(self, m: ShiftedWindowAttention, weight_bits=8, input_bits=8, aq_learnable=True, wq_learnable=True,
                 symmetric=True, weight_channelwise=True, input_channelwise=True, weight_quant_method="statsq",
                 input_quant_method="lsq", pretrained_initialized=False, **kwargs):
        _check_window_attention_like(m)
        super().__init__(dim=m.dim, window_size=m.window_size, shift_size=m.shift_size, num_heads=m.num_heads, qkv_bias=True,
                         proj_bias=True, attention_dropout=0.0, dropout=0.0, qqkkvv=getattr(m, "qqkkvv", False))
        if symmetric is False:
            raise NotImplementedError("unsigned input quantization of the shared attention input is not used by any recipe")
        self.weight_bits, self.input_bits, self.input_channelwise = weight_bits, input_bits, input_channelwise
        dim = m.qkv.in_features
        self.scale = (dim // m.num_heads) ** -0.5
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
        self.proj = QLinear(m=self.proj, **kw)
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

    def forward(self, x):
        return self._window_forward(x, lambda xw: xw)


class QAttention_swin_qkreparam_4_cga(QAttention_swin_qkreparam):
    """swin_attention_and_mlp.py:463-671: value- and gradient-identical to QAttention_swin_qkreparam."""

    def __init__(self, m: ShiftedWindowAttention, boundaryRange=0.005, **kwargs):
        self._boundaryRange = boundaryRange
        super().__init__(m, **kwargs)

    def _make_qk_quant(self, wq_learnable, kwargs):
        return StatsQuantizer_specific_4_qkreparam_cga(num_bits=self.weight_bits, clip_learnable=wq_learnable,
                                                       boundaryRange=self._boundaryRange)
