"""LSQ activation quantizers (reference: src/quantization/quantizer/lsq.py).

LsqQuantizer / LsqQuantizer4v are the 2/3/4-bit quantizers of the hot path: as stand-alone modules they run the
codes kernel + dequantise; inside the fused layers they only own the learned step size `s` (same state-dict key)
and its lazy, data-dependent initialisation (lsq.py:544-569, 730-754).
The 8-bit "ends" (patch-embed input / conv weight, head input / weight: lsq.py:20-109, 306-513) stay
torch-composed for now and reproduce the reference arithmetic op for op.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..functional import ImgLsqFn, LsqFn, levels


def _eff_scale(alpha, g):
    """grad_scale(clip(alpha, 1e-5), g) of lsq.py:6-18, 593 with the reference's exact value and gradient."""
    ac = torch.where(alpha > 1e-5, alpha, 1e-5)      # scalar operands: no host->device copy (CUDA-graph safe)
    ac = alpha - alpha.detach() + ac.detach()
    ag = ac * g
    return (ac - ag).detach() + ag


def _round_pass(x):
    r = x.round()
    return (r - x).detach() + x


class _LsqBase(nn.Module):
    def __init__(self, bit, all_positive=False, per_channel=True, learnable=True, **kwargs):
        super().__init__()
        if bit == 1:
            self.thd_neg, self.thd_pos = (0, 1) if all_positive else (-1, 1)
        else:
            self.thd_neg, self.thd_pos = levels(bit, all_positive)
        self.bit = bit
        self.per_channel = per_channel
        self.all_positive = all_positive
        self.learnable = learnable
        self.register_parameter("s", None)
        self.initialized_alpha = False

    def _set_s(self, init_val):
        self.s = nn.Parameter(init_val.detach().clone().to(torch.float32), requires_grad=bool(self.learnable))
        self.initialized_alpha = True

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # `s` is created lazily (lsq.py:541); accept it from a checkpoint even before the first forward
        key = prefix + "s"
        if key in state_dict and self.s is None:
            self._set_s(state_dict[key])
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def extra_repr(self):
        return (f"bit={self.bit}, all_positive={self.all_positive}, s_learnable={self.learnable}, "
                f"per_channel={self.per_channel}")


class LsqQuantizer(_LsqBase):
    """lsq.py:515-610: one step size per index of dim -2 (token / query row / (token, head))."""

    def init_from(self, x, *args, **kwargs):
        f = 4 if self.all_positive else 2
        a = x.detach().abs().mean(dim=-1)
        if x.dim() == 3:
            a = a.mean(dim=0)
        elif x.dim() == 4:
            a = a.mean(dim=0).mean(dim=0)
        self._set_s(f * a / (self.thd_pos ** 0.5))

    def forward(self, x):
        if not self.per_channel:
            raise NotImplementedError("per-tensor LsqQuantizer is not used by any OFQ recipe")
        if not self.initialized_alpha:
            self.init_from(x)
        return LsqFn.apply(x, self.s, self.bit, self.all_positive, False)


class LsqQuantizer4v(_LsqBase):
    """lsq.py:701-800: one step size per last-dim channel."""

    def init_from(self, x, *args, **kwargs):
        f = 4 if self.all_positive else 2
        a = x.detach().abs()
        for _ in range(x.dim() - 1):
            a = a.mean(dim=0)
        self._set_s(f * a / (self.thd_pos ** 0.5))

    def forward(self, x):
        if not self.initialized_alpha:
            self.init_from(x)
        return LsqFn.apply(x, self.s, self.bit, self.all_positive, True)


# ------------------------------------------------------------------------------------------------ 8-bit ends
class _Lsq8(_LsqBase):
    def _quant(self, x, alpha, count):
        g = 1.0 / ((self.thd_pos * count) ** 0.5)
        s = _eff_scale(alpha, g)
        x = x / s
        x = torch.clamp(x, self.thd_neg, self.thd_pos)
        return _round_pass(x) * s


class LsqQuantizerWeight(_Lsq8):
    """lsq.py:20-109: per-output-row learned step size of the 8-bit head weight."""

    def __init__(self, bit=8, all_positive=False, per_channel=True, learnable=True, **kwargs):
        if not per_channel or all_positive:
            # the reference also has a scalar-step form and a 4x unsigned init (lsq.py:72-101); no OFQ recipe uses them
            raise NotImplementedError("LsqQuantizerWeight: only the per-row signed form (per_channel=True, all_positive=False) is built")
        super().__init__(bit, all_positive, per_channel, learnable)

    def forward(self, x):
        if not self.initialized_alpha:
            self._set_s(2 * x.detach().abs().mean(dim=-1) / (self.thd_pos ** 0.5))
        return self._quant(x, self.s.unsqueeze(-1), x.shape[-1])


class LsqQuantizer4head_input(_Lsq8):
    """lsq.py:448-513: scalar step size of the 8-bit head input."""

    def forward(self, x):
        if not self.initialized_alpha:
            self._set_s((x.detach().abs().mean() * 2 / (self.thd_pos ** 0.5)).reshape(1))
        return self._quant(x, self.s, x.numel())


class LsqQuantizer4Conv2d(_Lsq8):
    """lsq.py:384-446: per-output-channel step size of the 8-bit patch-embed conv weight."""

    def __init__(self, bit=8, all_positive=False, per_channel=True, learnable=True, **kwargs):
        if not per_channel or all_positive:
            raise NotImplementedError("LsqQuantizer4Conv2d: only the per-output-channel signed form is built (lsq.py:419-437)")
        super().__init__(bit, False, per_channel, learnable)

    def forward(self, x):
        if not self.initialized_alpha:
            self._set_s(2 * x.detach().abs().mean(dim=-1).mean(dim=-1).mean(dim=-1) / (self.thd_pos ** 0.5))
        return self._quant(x, self.s.view(-1, 1, 1, 1), x.shape[1] * x.shape[2] * x.shape[3])


class LsqQuantizer4img(_Lsq8):
    """lsq.py:306-382: per-input-channel step size of the 8-bit image quantizer; the sign of the data is
    detected on the fly and latched in the `signed` buffer (lsq.py:338-341). The latch is read back to the host
    only until it is set or the scale is initialised (the reference syncs every forward)."""

    def __init__(self, bit=8, all_positive=False, per_channel=True, learnable=True, **kwargs):
        super().__init__(bit, all_positive, per_channel, learnable)
        self.register_buffer("signed", torch.zeros(1))
        self._signed_host = None

    def forward(self, x):
        if self._signed_host is None or (self._signed_host == 0):
            if float(self.signed) != 0 or float(x.detach().min()) < -1e-5:
                self.signed.data.fill_(1)
                self._signed_host = 1
            else:
                self._signed_host = 0
        if self._signed_host == 0:
            self.thd_neg, self.thd_pos = 0, 2 ** self.bit - 1
        else:
            self.thd_neg, self.thd_pos = -(2 ** (self.bit - 1)), 2 ** (self.bit - 1) - 1
        if not self.initialized_alpha:
            f = 4 if self.all_positive else 2
            self._set_s(f * x.detach().abs().mean(dim=-1).mean(dim=-1).mean(dim=0) / (self.thd_pos ** 0.5))
        return self._quant(x, self.s.view(1, -1, 1, 1), x.shape[0] * x.shape[2] * x.shape[3])

    def fused_ready(self, x, move_b4, move_aft) -> bool:
        """The step size exists, the data is known to be signed (8-bit codes in [-128, 127]) and the image is an fp32 CUDA
        batch whose pixels match the per-pixel shifts: the hot-path LSQ kernels apply in the [B*Cin, H*W] view."""
        if x.dim() != 4:
            return False
        H, W = x.shape[-2], x.shape[-1]
        return bool(self.initialized_alpha and self._signed_host == 1 and x.is_cuda and x.dtype == torch.float32
                    and H == W and move_b4.bias.numel() == H * W and move_aft.bias.numel() == H * W and (H * W) % 4 == 0
                    and self.bit <= 8 and x.shape[1] == self.s.numel())

    def forward_with_shifts(self, x, move_b4, move_aft):
        """move_aft(self(move_b4(x))) (qlinear.py:171). Once the step size exists and the data is known to be signed (8-bit
        codes in [-128, 127]), fp32 CUDA images take ONE fused kernel per direction (functional.ImgLsqFn)."""
        if self.fused_ready(x, move_b4, move_aft):
            return ImgLsqFn.apply(x, move_b4.bias, move_aft.bias, self.s, -(2 ** (self.bit - 1)), 2 ** (self.bit - 1) - 1)
        return move_aft(self(move_b4(x)))
