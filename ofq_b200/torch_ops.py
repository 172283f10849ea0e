"""The hot-path kernels as PyTorch dispatcher operators, `torch.ops.ofq_b200.*` ("ops registered with
TORCH_LIBRARY(ofq_b200, ...), each a thin wrapper over an extern "C" launcher").

The registration is done from Python with `torch.library` (the same dispatcher tables TORCH_LIBRARY fills; the launchers stay
the C-ABI of `include/ofq_b200.h`, reached through `ofq_b200.ops`). Each operator has a CUDA implementation only - calling one
with CPU tensors raises, as the product path has no CPU fallback - and a fake (meta) implementation, so the operators trace
under `torch.compile` / `torch.export` / FakeTensor shape propagation. Import this module to register them:

    import ofq_b200.torch_ops
    codes, colscale = torch.ops.ofq_b200.statsq_codes(w, 2)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops

_NS = "ofq_b200"


def _op(name, mutates=()):
    return torch.library.custom_op(f"{_NS}::{name}", mutates_args=mutates, device_types="cuda")


# ------------------------------------------------------------------------------------------------ quantizers
@_op("statsq_codes")
def statsq_codes(w: Tensor, bits: int) -> Tuple[Tensor, Tensor]:
    """StatsQuantizer.forward as integer codes (statsq.py:133-150): (int8 codes 2k+1 [R, C], fp32 colscale [R])."""
    codes, colscale = ops.statsq_codes(w.contiguous(), bits)[:2]
    return codes.clone(), colscale.clone()          # (the step prologue may own the buffers: hand out copies)


@statsq_codes.register_fake
def _(w, bits):
    return w.new_empty(w.shape, dtype=torch.int8), w.new_empty(w.shape[0], dtype=torch.float32)


@_op("lsq_effective_scale")
def lsq_effective_scale(alpha: Tensor, g: float) -> Tensor:
    """grad_scale(clip(alpha, 1e-5), g) value (lsq.py:6-18, 593), bit-exact."""
    return ops.lsq_effective_scale(alpha, g).clone()


@lsq_effective_scale.register_fake
def _(alpha, g):
    return torch.empty_like(alpha)


@_op("lsq_quant")
def lsq_quant(x: Tensor, b4: Tensor, s_eff: Tensor, per_col: bool, period: int, nseg: int, lo: int, hi: int) -> Tensor:
    """LearnableBias + LsqQuantizer(.4v) forward as int8 codes rint(clamp((x + b4) / s, lo, hi)) (lsq.py:595-601, 784-788);
    x [rows, cols]; per-row scales index (row % period) * nseg + col // (cols / nseg), per-column scales index col."""
    return ops.lsq_quant(x, b4, s_eff, ops.PER_COL if per_col else ops.PER_ROW, period, nseg, lo, hi)


@lsq_quant.register_fake
def _(x, b4, s_eff, per_col, period, nseg, lo, hi):
    return x.new_empty(x.shape, dtype=torch.int8)


# ------------------------------------------------------------------------------------------------ GEMMs
@_op("qgemm_fwd")
def qgemm_fwd(a_codes: Tensor, a_rowscale: Tensor, w_codes: Tensor, w_colscale: Tensor, colterm: Optional[Tensor]) -> Tensor:
    """out[m, n] = (sum_k a_codes[m, k] w_codes[n, k]) * a_rowscale[m % len] * w_colscale[n] + colterm[n]: the exact int8
    tensor-core GEMM of QLinear (qlinear.py:69) with every scale in the epilogue."""
    M, K = a_codes.shape
    N = w_codes.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a_codes.device)
    ops.gemm(ops.GEMM_I8, a_codes, (a_codes.stride(0), 0, 0, 0), w_codes, (w_codes.stride(0), 0, 0, 0), out, (N, 0, 0), M, N, K,
             rs=ops.vec(a_rowscale, a_rowscale.numel()), cs=ops.vec(w_colscale), ct=ops.vec(colterm) if colterm is not None else None)
    return out


@qgemm_fwd.register_fake
def _(a_codes, a_rowscale, w_codes, w_colscale, colterm):
    return a_codes.new_empty((a_codes.shape[0], w_codes.shape[0]), dtype=torch.float32)


@_op("qgemm_lsq_fwd")
def qgemm_lsq_fwd(a_codes: Tensor, a_rowscale: Tensor, w_codes: Tensor, w_colscale: Tensor, colterm: Optional[Tensor], b4: Tensor,
                  s2: Tensor, period: int, nseg: int, lo: int, hi: int) -> Tensor:
    """qgemm_fwd with the NEXT quantizer in its epilogue (ofq_gemm_lsq): int8 codes of LSQ(out + b4); s2 = [s_eff, 1 / s_eff]."""
    M, K = a_codes.shape
    N = w_codes.shape[0]
    return ops.gemm_lsq(a_codes, w_codes, M, N, K, b4, s2, period, nseg, lo, hi, rs=ops.vec(a_rowscale, a_rowscale.numel()),
                        cs=ops.vec(w_colscale), ct=ops.vec(colterm) if colterm is not None else None)[0]


@qgemm_lsq_fwd.register_fake
def _(a_codes, a_rowscale, w_codes, w_colscale, colterm, b4, s2, period, nseg, lo, hi):
    return a_codes.new_empty((a_codes.shape[0], w_codes.shape[0]), dtype=torch.int8)


@_op("wqk_compose")
def wqk_compose(wq: Tensor, wk: Tensor, num_heads: int) -> Tensor:
    """Per-head W_q[h]^T W_k[h] in fp32 (attention.py:190-194): [H * C, C]."""
    return ops.wqk_compose(wq.contiguous(), wk.contiguous(), num_heads).clone()


@wqk_compose.register_fake
def _(wq, wk, num_heads):
    return wq.new_empty((num_heads * wq.shape[1], wq.shape[1]))


# ------------------------------------------------------------------------------------------------ CGA / optimizer
@_op("cga_mask")
def cga_mask(w: Tensor, bits: int, boundary_range: float) -> Tensor:
    """freeze_outside_boundary_weight_idx (cga.py:450-469): uint8, 1 = frozen."""
    return ops.cga_mask(w.contiguous(), bits, boundary_range)


@cga_mask.register_fake
def _(w, bits, boundary_range):
    return w.new_empty(w.shape, dtype=torch.uint8)


@_op("cga_adamw_step", mutates=("p", "exp_avg", "exp_avg_sq"))
def cga_adamw_step(p: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, step: int, lr: float, beta1: float, beta2: float,
                   eps: float, weight_decay: float, bits: int, boundary_range: float) -> None:
    """One (CGA-masked when bits > 0) AdamW step in place (cga.py:953-1013 around torch.optim.AdamW)."""
    ops.cga_adamw_(p, grad.contiguous(), exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, weight_decay, bits=bits,
                   boundary_range=boundary_range)


OPS = ("statsq_codes", "lsq_effective_scale", "lsq_quant", "qgemm_fwd", "qgemm_lsq_fwd", "wqk_compose", "cga_mask", "cga_adamw_step")
