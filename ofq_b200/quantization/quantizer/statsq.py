"""StatsQuantizer modules (reference: src/quantization/quantizer/statsq.py:122-193)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops


class _StatsQFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, bits):
        codes, colscale, sf, _, _ = ops.statsq_codes(weight.contiguous(), bits, want_sf=True)
        ctx.mark_non_differentiable(sf)
        return codes.to(torch.float32) * colscale.unsqueeze(1), sf

    @staticmethod
    def backward(ctx, g, _gsf):
        return g, None            # statsq.py:148: identity straight-through for every element


class StatsQuantizer(nn.Module):
    """statsq.py:122-150. forward(weight[out, in]) -> fake-quantized weight; identity STE gradient.
    In the fused layers (QLinear, QAttention*) this module only owns `clip_val` for state-dict parity; the
    layers call the codes kernel directly."""

    def __init__(self, num_bits, clip_learnable):
        super().__init__()
        self.num_bits = num_bits
        self.clip_val = nn.Parameter(torch.Tensor([2.0]), requires_grad=False)
        self._s_dev = None

    @property
    def s(self):
        """Per-channel scale of the last forward (statsq.py:143), fetched lazily: no per-forward host sync."""
        return None if self._s_dev is None else self._s_dev.detach().cpu()

    @s.setter
    def s(self, value):
        self._s_dev = value

    def forward(self, weight):
        if weight.dim() != 2:
            raise ValueError("StatsQuantizer (B200): only 2-D weights are on the OFQ hot path")
        out, sf = _StatsQFn.apply(weight, self.num_bits)
        self._s_dev = sf
        return out


class StatsQuantizer_specific_4_qkreparam_cga(StatsQuantizer):
    """statsq.py:154-193. Value- and gradient-identical to StatsQuantizer: the boundary-band mixing at
    statsq.py:181-188 only feeds a detached tensor (verified against the reference, tests/golden/statsq.npz)."""

    def __init__(self, num_bits, clip_learnable, boundaryRange=0.005):
        super().__init__(num_bits, clip_learnable)
        self.boundaryRange = boundaryRange
