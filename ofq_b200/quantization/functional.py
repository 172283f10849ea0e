"""Autograd functions of the OFQ hot path: each forward/backward is a fixed sequence of sm_100a kernels
reached through the C-ABI (ofq_b200/ops.py).  No arithmetic of the path is done by PyTorch here; torch only
allocates buffers and re-lays-out a few per-token scale vectors (hundreds of floats).

Data layout (HBM): activations are fp32 [B*N, C] row-major; every quantized operand is stored ONCE as int8
codes (+ an fp32 scale vector and the `move_aft` shift), so the forward GEMMs are exact integer GEMMs and the
backward rebuilds bf16 operands from the saved codes.  Reference: src/quantization/modules/{qlinear,attention}.py.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from .. import ops, prologue
from ..ops import ACT_GELU, ACT_NONE, FMT_BF16, FMT_F16, GEMM_BF16, GEMM_F16, GEMM_I8, PER_COL, PER_ROW, round_up, vec



def _num_sms() -> int:
    """SM count of the current device (split-K heuristic of the bf16 modes; the fp16 default lets the library choose)."""
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count

# Number format of the real-valued (gradient) operand of every backward GEMM; the integer-code operand is exact in all:
#   "f16"    one fp16 plane (11 significant bits, ~2e-4 gradient error), range-scaled per tensor by a power of two
#            from an absmax pass (ops.absmax_scale) and un-scaled in the GEMM epilogue              [default]
#   "bf16x2" bf16 hi + lo planes (~16 significant bits, ~5e-6 gradient error, twice the tensor work)
#   "bf16"   one bf16 plane (1-2e-3 gradient error: outside the 1e-3 parity bound, kept for experiments)
_mode = os.environ.get("OFQ_BWD_MODE")
if _mode is None:
    _mode = {"2": "bf16x2", "1": "bf16"}.get(os.environ.get("OFQ_BWD_PLANES", ""), "f16")
if _mode not in ("f16", "bf16x2", "bf16"):
    raise ValueError(f"OFQ_BWD_MODE={_mode!r}: expected f16, bf16x2 or bf16")
BWD_MODE = _mode
F16 = BWD_MODE == "f16"
PLANES = 2 if BWD_MODE == "bf16x2" else 1
GEMM_BWD = GEMM_F16 if F16 else GEMM_BF16
FMT = FMT_F16 if F16 else FMT_BF16
# With two planes the GEMM loads hi and lo of the gradient operand in the same pipeline stage and multiplies both by
# one copy of the code operand ("dual-A", ofq_operand_t.dual_delta) instead of looping over planes as outer-K slices.
DUAL = PLANES == 2
K2P = 1 if DUAL else PLANES          # outer-K slices still looped over
DD = 1 if DUAL else 0                # dual_delta for operands laid out [plane][...]
# fp16 mode: let the LSQ backward of the V / qkx quantizers write the fp16 GEMM operand directly (tests toggle this)
FUSED16 = os.environ.get("OFQ_FUSED16", "1") != "0"
# Fused attention forward (ofq_qkr_attn_fwd: scores -> softmax -> probability codes -> P.V in one kernel, logits never in HBM)
# for the QKR path wherever its shape limits hold (head dim 64, <= 208 tokens, C <= 384: DeiT-T / DeiT-S); OFQ_FUSED_ATTN=0
# keeps the three-kernel path (A/B measurements, bit-identity tests).
FUSED_ATTN = os.environ.get("OFQ_FUSED_ATTN", "1") != "0"
# the qkx quantizer as the epilogue of the qkx GEMM (ofq_gemm_lsq); 0 = GEMM -> fp32 qkx -> ofq_lsq_quant (A/B measurements, tests)
FUSED_QKX = os.environ.get("OFQ_FUSED_QKX", "1") != "0"
# the input quantizer's backward as the epilogue of the layer's dX GEMM (ofq_gemm_dx_lsq); 0 = GEMM -> fp32 dX_hat -> ofq_lsq_bwd
FUSED_DX = os.environ.get("OFQ_FUSED_DX", "1") != "0"
# ... and its backward (ofq_qkr_attn_bwd: logits recomputed, dP = dO v^T, softmax / quantizer backward in one kernel; neither P
# nor dP in HBM). fp16 mode only; OFQ_FUSED_ATTN_BWD=0 keeps dP GEMM + ofq_softmax_quant_bwd on the saved probabilities.
FUSED_ATTN_BWD = os.environ.get("OFQ_FUSED_ATTN_BWD", "1") != "0"
# Debug tap (parity tests): when a list, every forward of the autograd functions below appends (kind, {name: tensor}) with
# the integer codes and the pre-quantizer values of each of its quantizers, in call order. None in production.
TAP = None


def levels(bit: int, all_positive: bool):
    """(thd_neg, thd_pos) of lsq.py:519-534."""
    if bit == 1:
        raise NotImplementedError("1-bit (sign) LSQ is not part of the OFQ recipes (2/3/4 bit)")
    if all_positive:
        return 0, 2 ** bit - 1
    return -(2 ** (bit - 1)), 2 ** (bit - 1) - 1


_GRAD_AT_APPLY = True


class _Fn(torch.autograd.Function):
    """autograd.Function whose forward can tell whether a backward can ever follow. Inside `forward` grad mode is always off, and
    `ctx.needs_input_grad` mirrors `requires_grad` of the inputs even under torch.no_grad(): neither says whether the caller is
    building a graph. `apply` records the caller's grad mode for `_need_grad`."""

    @classmethod
    def apply(cls, *args, **kwargs):
        global _GRAD_AT_APPLY
        prev = _GRAD_AT_APPLY
        _GRAD_AT_APPLY = torch.is_grad_enabled()
        try:
            return super().apply(*args, **kwargs)
        finally:
            _GRAD_AT_APPLY = prev


def _need_grad(ctx) -> bool:
    """Will this node ever run a backward? (Under no_grad inference must not produce the fp16 operand copies, probabilities and
    residual planes only a backward reads.)"""
    return _GRAD_AT_APPLY and any(ctx.needs_input_grad)


def grad_scale_factor(hi: int, count: int) -> float:
    """s_grad_scale of lsq.py:582-591 / 775-778: 1/sqrt(thd_pos * elements-per-scale)."""
    return 1.0 / ((hi * count) ** 0.5)


def _splits_for(tiles: int, kblocks: int) -> int:
    s = max(1, min((2 * _num_sms() + tiles - 1) // tiles, max(1, kblocks // 4)))
    return min(s, 64)


def _scalar(sc):
    """The un-scale 1/sc of a range-scaled fp16 operand as a period-1 epilogue vector."""
    return vec(sc[1:2], 1)


# Set by ddp.FlatGradAllReduce.zero() for the duration of a step: `.take(weight)` hands out the weight's (zeroed) slice of
# the flat gradient buffer, so the dW GEMM accumulates in place and no gather copy / per-weight fill is needed.
GRAD_SLOTS = None


def _dw_buffer(w, Nout: int, K: int, device) -> torch.Tensor:
    if GRAD_SLOTS is not None and w is not None and tuple(w.shape) == (Nout, K):
        v = GRAD_SLOTS.take(w)
        if v is not None:
            return v
    return torch.zeros((Nout, K), dtype=torch.float32, device=device)


def _linear_backward_f16(dY2d, qx, wc, cs2, se2, period, x_aft, dxhat, accumulate_dx: bool, qx16=None, sc=None, a16=None,
                         colsum=None, amax_dx=None, wc16=None, dw_for=None, dx_lsq=None):
    """fp16 backward of out = x_hat @ W_hat^T (+bias) with ONE range-scaled copy of the gradient,
    A16[t,n] = fp16(dY[t,n] * colscale[n] * se_x[t] * sc), read K-major by the dX GEMM and MN-major by the dW GEMM; the
    code operands stay exact and un-transposed (MN-major B), the folded scale vectors are undone per output row:
        dX_hat[t,k] (+)= 1/(se_x[t] sc) * sum_n A16[t,n] wc[n,k]
        dW[n,k]        = 1/(colscale[n] sc) * sum_t A16[t,n] qx[t,k] + colsum(dY)[n] * aft[k]
    cs2 = [colscale, 1/colscale], se2 = [se_x, 1/se_x]. Returns (dW, dbias, qx16).
    a16 / colsum: the operand and colsum(dY) already produced by the pass that made dY (ops.lsq_bwd(out16=...)); dY2d is
    then not read (and may be None)."""
    M, K = qx.shape
    Nout = wc.shape[0]
    if a16 is None:
        if sc is None:
            bmax = ops.residual_max_for(dY2d)       # dY straight out of an ofq_b200 LayerNorm backward: maxima already known
            if bmax is not None:
                sc = ops.scale_from_max(bmax, v1=cs2[0], v2=se2[0], product=True)
        if sc is None:
            sc = ops.absmax_scale(dY2d, 1, M, Nout, dY2d.stride(0), 0, cs=cs2[0], rs=se2[0], rs_period=period, product=True)
        prep = ops.grad_prep(dY2d, 1, M, Nout, dY2d.stride(0), 0, cs=cs2[0], rs=se2[0], rs_period=period, want_rm=True,
                             want_colsum=True, fmt=FMT, scale4=sc, rm_rowscale=True)
        a16, colsum = prep["rm"], prep["colsum"]
    if wc16 is None:
        wc16 = ops.codes_to_bf16(wc, 1, Nout, K, K, 0, False, FMT)       # [1, Nout, K]
    lsq_grads = None
    if dx_lsq is not None:
        # the dX GEMM with the input quantizer's backward (STE mask, ds, db4, daft) as its epilogue: dX_hat is never written
        x2d_, b4_, lo_, hi_, g_, amax_ = dx_lsq
        lsq_grads = ops.gemm_dx_lsq(GEMM_BWD, a16, (Nout, 0, 0, 0), wc16, (K, 0, 0, 0), M, K, Nout, rs=vec(se2[1], period), cs=_scalar(sc),
                                    x2d=x2d_, b4=b4_, period=period, qlo=lo_, qhi=hi_, g=g_, w_codes=wc, dy_colsum=colsum,
                                    colscale=cs2[0], b_mn=True, amax=amax_)
    else:
        ops.gemm(GEMM_BWD, a16, (Nout, 0, 0, 0), wc16, (K, 0, 0, 0), dxhat, (K, 0, 0), M, K, Nout, b_mn=True,
                 accumulate=accumulate_dx, rs=vec(se2[1], period), cs=_scalar(sc), amax=amax_dx)
    if qx16 is None:
        qx16 = ops.codes_to_bf16(qx, 1, M, K, K, 0, False, FMT)          # [1, M, K]
    dW = _dw_buffer(dw_for, Nout, K, a16.device)
    # the weight gradient is off the activation-gradient chain: side stream, joined before the backward returns
    with ops.side_stream(a16, qx16, dW, cs2, sc, colsum, x_aft):
        ops.gemm(GEMM_BWD, a16, (Nout, 0, 0, 0), qx16, (K, 0, 0, 0), dW, (K, 0, 0), Nout, K, M, a_mn=True, b_mn=True,
                 splits=0, accumulate=True, rs=vec(cs2[1]), cs=_scalar(sc), rt=vec(colsum), ct=vec(x_aft))
    if lsq_grads is not None:
        return dW, colsum, qx16, lsq_grads
    return dW, colsum, qx16


def _linear_backward(dY2d, qx, wc, cs2, se2, period, x_aft, dxhat, accumulate_dx: bool, qxT_all=None, sc=None, **f16kw):
    """Backward of out = x_hat @ W_hat^T (+bias):  dX_hat (+)= dY W_hat,  dW = dY^T x_hat,  dbias = colsum(dY).
    x_hat = qx * se_x[row % period] + x_aft,  W_hat = wc * colscale[row].  Returns (dW, dbias, shared code operand)."""
    if F16:
        return _linear_backward_f16(dY2d, qx, wc, cs2, se2, period, x_aft, dxhat, accumulate_dx, qxT_all, sc, **f16kw)
    colscale, se_x = cs2[0], se2[0]
    M, Nout = dY2d.shape
    K = qx.shape[1]
    prep = ops.grad_prep(dY2d, 1, M, Nout, dY2d.stride(0), 0, cs=colscale, rs=se_x, rs_period=period,
                         want_rm=True, want_t=True, want_colsum=True, planes=PLANES)
    m_pad = prep["r_pad"]
    # dX_hat[M,K] = (dY * colscale)[M,Nout] @ codes[Nout,K]
    wcT = ops.codes_to_bf16(wc, 1, Nout, K, K, 0, True, FMT)       # [1, K, nout_pad]
    nout_pad = wcT.shape[-1]
    ops.gemm(GEMM_BWD, prep["rm"], (Nout, M * Nout, 0, 0), wcT, (nout_pad, 0, 0, 0), dxhat, (K, 0, 0), M, K, Nout,
             k2=K2P, a_dual_delta=DD, accumulate=accumulate_dx)
    # dW[Nout,K] = (dY * se_x)^T[Nout,M] @ qx[M,K]  + colsum(dY)[Nout] x aft[K]
    if qxT_all is None:
        qxT_all = ops.codes_to_bf16(qx, 1, M, K, K, 0, True, FMT)  # [1, K, m_pad]
    dW = _dw_buffer(f16kw.get("dw_for"), Nout, K, dY2d.device)
    tiles = ((Nout + 127) // 128) * ((K + 127) // 128)
    splits = _splits_for(tiles, K2P * ((M + 63) // 64))
    ops.gemm(GEMM_BWD, prep["t"], (m_pad, Nout * m_pad, 0, 0), qxT_all, (m_pad, 0, 0, 0), dW, (K, 0, 0), Nout, K, M,
             k2=K2P, a_dual_delta=DD, splits=splits, accumulate=True, rt=vec(prep["colsum"]), ct=vec(x_aft))
    return dW, prep["colsum"], qxT_all


_ZERO = {}


def _placeholder_grad(like: torch.Tensor) -> torch.Tensor:
    """A gradient-shaped, zero-stride view of ONE cached zero: what a node returns for an input whose real gradient
    travels another way (MlpLink.fuse). It costs no memory and no launch, reads as exact zeros (never as uninitialised
    memory) and is recognisable by its data pointer."""
    z = _ZERO.get(like.device)
    if z is None:
        z = _ZERO[like.device] = torch.zeros((), dtype=like.dtype, device=like.device)
    return z.expand(like.shape)


def _is_placeholder_grad(t: torch.Tensor) -> bool:
    z = _ZERO.get(t.device)
    return z is not None and t.data_ptr() == z.data_ptr() and all(s == 0 for s in t.stride())


# ====================================================================================== QLinear
class MlpLink:
    """Side channel between a producer node and the QLinearFn that consumes its output (fc1 -> fc2 of a QMLP, attention
    core -> proj): the producer's forward leaves the two scale vectors folded into its gradient operand here (cs per
    column, se per row), the backward of the consumer (which runs first) turns the max |d producer_out| its LSQ pass sees
    anyway into the fp16 range scale of that operand (`sc`), and the producer's backward then skips its absmax pass."""
    __slots__ = ("cs", "se", "sc", "fuse", "a16", "colsum")

    def __init__(self, fuse: bool = False):
        # fuse: the consumer's LSQ backward writes the producer's fp16 gradient operand (a16) and colsum(dY) itself; the
        # fp32 gradient handed back through autograd is then a zero placeholder (_placeholder_grad) that the producer
        # recognises and ignores. Anything else that consumed the producer's output (a hook, a second use) would add its
        # own gradient to the placeholder: the producer detects that and fails loudly instead of dropping it
        self.cs = self.se = self.sc = self.a16 = self.colsum = None
        self.fuse = fuse


class QLinearFn(_Fn):
    """QLinear.forward (qlinear.py:58-73): StatsQ weight codes, (move_b4 -> LSQ -> move_aft) input codes,
    int8 tcgen05 GEMM with the scales / shift / bias in the epilogue.
    act = ACT_GELU: the layer computes QLinear(GELU(x)) (fc2 of QMLP, qlinear.py:123-136) with the activation fused into
    the quantizer pass; x is then the saved fc1 output and the returned gradient is w.r.t. that pre-activation.
    link / role: optional MlpLink shared by fc1 (role 1) and fc2 (role 2) of one QMLP."""

    @staticmethod
    def forward(ctx, x, weight, bias, b4, aft, s, wbits: int, abits: int, unsigned: bool, act: int = ACT_NONE,
                link: Optional[MlpLink] = None, role: int = 0):
        K = x.shape[-1]
        P = x.shape[-2]
        xc = x.contiguous()
        x2d = xc.view(-1, K)
        M = x2d.shape[0]
        Nout = weight.shape[0]
        lo, hi = levels(abits, unsigned)
        g = grad_scale_factor(hi, x.numel() // P)
        se2 = ops.lsq_effective_scale(s, g, recip=True)
        se = se2[0]
        qx16 = None
        if F16 and _need_grad(ctx):
            qx, qx16 = ops.lsq_quant(x2d, b4, se, PER_ROW, P, 1, lo, hi, act=act, fmt16=FMT)
        else:
            qx = ops.lsq_quant(x2d, b4, se, PER_ROW, P, 1, lo, hi, act=act)
        w16 = FMT if (F16 and _need_grad(ctx)) else None
        wc, colscale, _, colterm, _, inv_cs, *wc16 = ops.statsq_codes(weight, wbits, aft=aft, bias=bias, want_inv=True, fmt16=w16)
        wc16 = wc16[0] if wc16 else None
        out = torch.empty((M, Nout), dtype=torch.float32, device=x.device)
        ops.gemm(GEMM_I8, qx, (K, 0, 0, 0), wc, (K, 0, 0, 0), out, (Nout, 0, 0), M, Nout, K,
                 rs=vec(se, P), cs=vec(colscale), ct=vec(colterm))
        if link is not None and role == 1:
            link.cs, link.se, link.sc = colscale, se, None
        if TAP is not None:
            TAP.append(("qlinear", dict(x=x2d, b4=b4, se=se, period=P, act=act, qx=qx, wc=wc, colscale=colscale, out=out)))
        ctx.save_for_backward(xc, qx, wc, colscale, inv_cs, se2, b4, aft, qx16, wc16, weight)
        ctx.cfg = (P, lo, hi, g, bias is not None, act, link, role)
        ctx.pro_gen = prologue.current_generation()
        return out.view(*x.shape[:-1], Nout)

    @staticmethod
    def backward(ctx, dY):
        prologue.check_generation(ctx.pro_gen)
        xc, qx, wc, colscale, inv_cs, se2, b4, aft, qx16, wc16, weight = ctx.saved_tensors
        P, lo, hi, g, has_bias, act, link, role = ctx.cfg
        wkw = {"wc16": wc16} if (F16 and wc16 is not None) else {}
        wkw["dw_for"] = weight
        K = xc.shape[-1]
        x2d = xc.view(-1, K)
        M = x2d.shape[0]
        fused_in = link is not None and role == 1 and link.a16 is not None
        # single-producer input gradients that leave as fp32 dx (fc1 and proj inputs): the quantizer's backward runs as the
        # epilogue of the dX GEMM (ofq_gemm_dx_lsq) and dX_hat never exists
        fuse_dx = (F16 and FUSED_DX and act == ACT_NONE and K % 64 == 0 and M % P == 0 and x2d.stride(0) % 4 == 0
                   and not (link is not None and role == 2 and link.fuse))
        # a consumer of dx that takes a range-scaled fp16 operand (the attention backward behind proj) needs max |dx|
        want_max = fuse_dx and link is not None and role == 2 and link.cs is not None and link.cs.shape[0] == K
        dx_amax = ops.scratch_zeros(1, dY.device) if want_max else None
        dxkw = {"dx_lsq": (x2d, b4, lo, hi, g, dx_amax)} if fuse_dx else {}
        dxhat = None if fuse_dx else torch.empty((M, K), dtype=torch.float32, device=dY.device)
        dY2d = None if fused_in else dY.contiguous().view(M, -1)      # (never materialise the zero-stride placeholder)
        if fused_in:
            # fc1 of a fused QMLP: fc2's backward already wrote this layer's fp16 gradient operand and colsum(dY); the dY
            # tensor that arrived through autograd is the zero placeholder
            if not _is_placeholder_grad(dY):
                raise RuntimeError("fused QMLP: the fc1 output received a gradient from something other than fc2 (a hook, "
                                   "retain_grad or a second consumer); run the MLP un-fused (OFQ_FUSED16=0) for such graphs")
            dW, dbias, _, *lsqg = _linear_backward_f16(None, qx, wc, (colscale, inv_cs), se2, P, aft, dxhat, False, qx16, sc=link.sc,
                                                       a16=link.a16, colsum=link.colsum, **wkw, **dxkw)
            link.a16 = link.colsum = None
        else:
            sc = link.sc if (link is not None and role == 1) else None
            fuse_next = (F16 and FUSED16 and link is not None and role == 2 and link.fuse and link.cs is not None
                         and link.cs.shape[0] == K and K % 4 == 0 and M % link.se.numel() == 0)
            amax = ops.scratch_zeros(1, dY.device) if fuse_next else None
            dW, dbias, _, *lsqg = _linear_backward(dY2d, qx, wc, (colscale, inv_cs), se2, P, aft, dxhat, False, qx16, sc=sc,
                                                   **({"amax_dx": amax} if fuse_next else {}), **wkw, **dxkw)
            if fuse_next:
                # the producer (fc1) only needs fp16(dx * colscale1[c] * se1[r] * sc) and colsum(dx): written by this pass
                # |dx| <= |dxhat| * max GELU' (1.13)
                sc1 = ops.scale_from_max(amax, v1=link.cs, v2=link.se, mult=1.13 if act == ACT_GELU else 1.0, product=True)
                _, ds, db4, daft, a16, csum = ops.lsq_bwd(dxhat, x2d, b4, se2[0], PER_ROW, P, 1, lo, hi, g, act=act,
                                                          out16=(FMT, link.cs, link.se, link.se.numel(), sc1), want_dx=False,
                                                          want_colsum=True)
                link.sc, link.a16, link.colsum = sc1, a16, csum
                dx = _placeholder_grad(xc)        # exact zeros, recognised by fc1's backward (see MlpLink.fuse)
                ops.side_join()
                return dx, dW, (dbias if has_bias else None), db4, daft, ds, None, None, None, None, None, None
        if lsqg:
            dx, ds, db4, daft = lsqg[0]
            if want_max:
                link.sc = ops.scale_from_max(dx_amax, v1=link.cs, v2=link.se, mult=1.0, product=True)
            ops.side_join()
            return dx.view_as(xc), dW, (dbias if has_bias else None), db4, daft, ds, None, None, None, None, None, None
        nxt = None
        if F16 and link is not None and role == 2 and link.cs is not None and link.cs.shape[0] == K:
            nxt = (link.cs, link.se, 1.0, True)
        dx, ds, db4, daft, *scn = ops.lsq_bwd(dxhat, x2d, b4, se2[0], PER_ROW, P, 1, lo, hi, g, act=act, next_scale=nxt)
        if nxt is not None:
            link.sc = scn[0]
        ops.side_join()
        return dx.view_as(xc), dW, (dbias if has_bias else None), db4, daft, ds, None, None, None, None, None, None


# ====================================================================================== standalone LSQ
class LsqFn(torch.autograd.Function):
    """LsqQuantizer / LsqQuantizer4v forward as a module of its own (lsq.py:571-602, 757-790)."""

    @staticmethod
    def forward(ctx, x, s, bit: int, all_positive: bool, per_col: bool):
        K = x.shape[-1]
        xc = x.contiguous()
        x2d = xc.view(-1, K)
        lo, hi = levels(bit, all_positive)
        if per_col:
            g = grad_scale_factor(hi, x.numel() // K)
            mode, period = PER_COL, 1
        else:
            period = x.shape[-2]
            g = grad_scale_factor(hi, x.numel() // period)
            mode = PER_ROW
        se = ops.lsq_effective_scale(s, g)
        zero = torch.zeros(K, dtype=torch.float32, device=x.device)
        codes = ops.lsq_quant(x2d, zero, se, mode, period, 1, lo, hi)
        # dequantise: q * s_eff  (exact product of a small integer and the scale, as the reference computes it)
        scale = se.view(1, -1) if per_col else se.repeat(x2d.shape[0] // period).view(-1, 1)
        out = (codes.to(torch.float32) * scale).view_as(xc)
        ctx.save_for_backward(xc, se, zero)
        ctx.cfg = (mode, period, lo, hi, g)
        return out

    @staticmethod
    def backward(ctx, dy):
        xc, se, zero = ctx.saved_tensors
        mode, period, lo, hi, g = ctx.cfg
        K = xc.shape[-1]
        dx, ds, _, _ = ops.lsq_bwd(dy.contiguous().view(-1, K), xc.view(-1, K), zero, se, mode, period, 1, lo, hi, g)
        return dx.view_as(xc), ds, None, None, None


class ImgLsqFn(torch.autograd.Function):
    """move_aft(LsqQuantizer4img(move_b4(img))) of the 8-bit patch-embedding input (qlinear.py:138-177, lsq.py:306-382,
    qbias.py:15-23) on the [B*Cin, H*W] view of the image: the per-input-channel step size is a per-row scale with period
    Cin, the per-pixel shifts are per-column vectors, so the hot-path kernels apply unchanged: one codes pass forward,
    one STE / reduction pass backward (the torch composition is ~20 elementwise passes over the batch of images)."""

    @staticmethod
    def forward(ctx, x, b4, aft, s, lo: int, hi: int):
        B, Cin, H, W = x.shape
        xc = x.contiguous()
        x2d = xc.view(B * Cin, H * W)
        g = grad_scale_factor(hi, B * H * W)
        se = ops.lsq_effective_scale(s, g)
        codes = ops.lsq_quant(x2d, b4, se, PER_ROW, Cin, 1, lo, hi)
        # dequantise exactly as the reference: round(.) * s, then the shift
        out = torch.mul(codes.view(B, Cin, H * W), se.view(1, Cin, 1)) + aft.view(1, 1, H * W)
        ctx.save_for_backward(xc, b4, se)
        ctx.cfg = (lo, hi, g)
        return out.view(B, Cin, H, W)

    @staticmethod
    def backward(ctx, dy):
        xc, b4, se = ctx.saved_tensors
        lo, hi, g = ctx.cfg
        B, Cin, H, W = xc.shape
        dx, ds, db4, daft = ops.lsq_bwd(dy.contiguous().view(B * Cin, H * W), xc.view(B * Cin, H * W), b4, se, PER_ROW, Cin, 1,
                                        lo, hi, g)
        return (dx.view_as(xc) if ctx.needs_input_grad[0] else None), db4, daft, ds, None, None


class HeadLinearFn(_Fn):
    """LSQ_QLinear4head.forward (qlinear.py:193-238), the 8-bit classifier heads, on the integer path:
    move_b4 -> LsqQuantizer4head_input (ONE learned step, lsq.py:448-513) -> move_aft on the input, LsqQuantizerWeight (one
    learned step per output row, lsq.py:20-109) on the weight, both as int8 codes, one exact int8 GEMM with the step sizes and
    the folded shift / bias term in its epilogue. Backward: the fp16 (or bf16x2) GEMM pair of every other linear layer, then the
    LSQ backward of both quantizers (straight-through masks, step-size gradients). Signed 8-bit only (codes must fit int8)."""

    @staticmethod
    def forward(ctx, x, weight, bias, b4, aft, s_x, s_w, bits: int):
        Nout, K = weight.shape
        xc = x.contiguous()
        x2d = xc.view(-1, K)
        M = x2d.shape[0]
        lo, hi = levels(bits, False)
        g_x = grad_scale_factor(hi, x2d.numel())
        g_w = grad_scale_factor(hi, K)
        sx2 = ops.lsq_effective_scale(s_x, g_x, recip=True)          # [2, 1]
        sw2 = ops.lsq_effective_scale(s_w, g_w, recip=True)          # [2, Nout]
        need_grad = _need_grad(ctx)
        f16 = FMT if (F16 and need_grad) else None
        qx16 = wc16 = None
        r = ops.lsq_quant(x2d, b4, sx2[0], PER_ROW, 1, 1, lo, hi, fmt16=f16)
        qx, qx16 = r if f16 is not None else (r, None)
        zero_k = ops.scratch_zeros(K, x.device)
        r = ops.lsq_quant(weight, zero_k, sw2[0], PER_ROW, Nout, 1, lo, hi, fmt16=f16)
        wc, wc16 = r if f16 is not None else (r, None)
        # out = se_x * se_w[n] * (qx . wc[n]) + se_w[n] * (aft . wc[n]) + bias[n]
        colterm = torch.addcmul(bias, ops.codes_rowdot(wc, 1, aft).view(-1), sw2[0])
        out = torch.empty((M, Nout), dtype=torch.float32, device=x.device)
        ops.gemm(GEMM_I8, qx, (K, 0, 0, 0), wc, (K, 0, 0, 0), out, (Nout, 0, 0), M, Nout, K, rs=vec(sx2[0], 1), cs=vec(sw2[0]), ct=vec(colterm))
        ctx.save_for_backward(xc, weight, b4, aft, qx, wc, sx2, sw2, qx16, wc16)
        ctx.cfg = (lo, hi, g_x, g_w)
        return out.view(*x.shape[:-1], Nout)

    @staticmethod
    def backward(ctx, dY):
        xc, weight, b4, aft, qx, wc, sx2, sw2, qx16, wc16 = ctx.saved_tensors
        lo, hi, g_x, g_w = ctx.cfg
        Nout, K = weight.shape
        x2d = xc.view(-1, K)
        M = x2d.shape[0]
        dY2d = dY.reshape(M, Nout)
        if not dY2d.is_contiguous():
            dY2d = dY2d.contiguous()
        dxhat = torch.empty((M, K), dtype=torch.float32, device=dY.device)
        dwhat, dbias, _ = _linear_backward(dY2d, qx, wc, (sw2[0], sw2[1]), sx2, 1, aft, dxhat, False, qx16,
                                           **({"wc16": wc16} if (F16 and wc16 is not None) else {}))
        dx, ds_x, db4, daft = ops.lsq_bwd(dxhat, x2d, b4, sx2[0], PER_ROW, 1, 1, lo, hi, g_x)
        zero_k = ops.scratch_zeros(K, dY.device)
        dW, ds_w, _, _ = ops.lsq_bwd(dwhat, weight, zero_k, sw2[0], PER_ROW, Nout, 1, lo, hi, g_w, want_aft=False)
        ops.side_join()
        return dx.view_as(xc), dW, dbias, db4, daft, ds_x, ds_w, None


class PatchEmbedFn(torch.autograd.Function):
    """The 8-bit patch-embedding convolution (qlinear.py:138-191) with stride = kernel as integer GEMMs on the tensor cores:
    move_b4 -> LsqQuantizer4img -> move_aft on the image (lsq.py:306-382, qbias.py:15-23), im2col of the int8 CODES, and

        out[b,p,n] = sum_c (s_c se_w[n]) sum_{ij} q[b,p,c,ij] qw[n,c,ij]  +  sum_{c,ij} aft[p,ij] W_hat[n,c,ij]  +  bias[n]

    one exact int8 GEMM per input channel (the per-channel step size s_c is constant inside it), accumulated in fp32 on top
    of the batch-independent shift term. `what` is the already fake-quantized weight (LsqQuantizer4Conv2d keeps its torch
    autograd: this node hands back d what), `se_w` its effective per-output-channel step (no gradient through it here).
    Backward (fp16 range-scaled operand like every other backward GEMM): d x_hat = col2im(dY W_hat) -> the LSQ backward of
    the image quantizer; d what = dY^T x_hat with the integer part on the tensor cores and the shift part as a small GEMM."""

    @staticmethod
    def forward(ctx, x, b4, aft, s_img, what, se_w, bias, lo: int, hi: int):
        B, Cin, H, W = x.shape
        Cout, _, kh, kw = what.shape
        gh, gw = H // kh, W // kw
        P, KK = gh * gw, kh * kw
        K, M = Cin * KK, B * P
        xc = x.contiguous()
        x2d = xc.view(B * Cin, H * W)
        g = grad_scale_factor(hi, B * H * W)
        se = ops.lsq_effective_scale(s_img, g)                       # [Cin]
        codes = ops.lsq_quant(x2d, b4, se, PER_ROW, Cin, 1, lo, hi)  # int8 [B*Cin, H*W]
        qcols = codes.view(B, Cin, gh, kh, gw, kw).permute(0, 2, 4, 1, 3, 5).contiguous().view(M, K)
        w2d = what.detach().reshape(Cout, K)
        qw = torch.round(w2d / se_w.view(-1, 1)).to(torch.int8)      # exact: what = code * se_w
        aftcols = aft.detach().view(gh, kh, gw, kw).permute(0, 2, 1, 3).reshape(P, KK)
        shift = torch.addmm(bias.detach(), aftcols, w2d.view(Cout, Cin, KK).sum(1).t())          # [P, Cout]
        out = shift.unsqueeze(0).expand(B, P, Cout).contiguous().view(M, Cout)
        csw = (se.view(-1, 1) * se_w.view(1, -1)).contiguous()       # [Cin, Cout]
        for c in range(Cin):
            ops.gemm(GEMM_I8, qcols[:, c * KK:], (K, 0, 0, 0), qw[:, c * KK:], (K, 0, 0, 0), out, (Cout, 0, 0), M, Cout, KK,
                     accumulate=True, cs=vec(csw[c]))
        ctx.save_for_backward(xc, b4, se, qcols, qw, se_w, aftcols)
        ctx.cfg = (lo, hi, g, B, Cin, H, W, Cout, kh, kw)
        return out.view(B, gh, gw, Cout).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        xc, b4, se, qcols, qw, se_w, aftcols = ctx.saved_tensors
        lo, hi, g, B, Cin, H, W, Cout, kh, kw = ctx.cfg
        gh, gw = H // kh, W // kw
        P, KK = gh * gw, kh * kw
        K, M = Cin * KK, B * P
        dY = dout.permute(0, 2, 3, 1).reshape(M, Cout)
        if not dY.is_contiguous():
            dY = dY.contiguous()
        # ONE fp16 copy A16[m,n] = fp16(dY se_w[n] sc), K-major for d x_hat and MN-major for d what
        sc = ops.absmax_scale(dY, 1, M, Cout, Cout, 0, cs=se_w)
        prep = ops.grad_prep(dY, 1, M, Cout, Cout, 0, cs=se_w, want_rm=True, want_colsum=True, fmt=FMT, scale4=sc)
        a16, dbias = prep["rm"], prep["colsum"]
        qw16 = ops.codes_to_bf16(qw, 1, Cout, K, K, 0, False, FMT)
        dxcols = torch.empty((M, K), dtype=torch.float32, device=dY.device)
        ops.gemm(GEMM_BWD, a16, (Cout, 0, 0, 0), qw16, (K, 0, 0, 0), dxcols, (K, 0, 0), M, K, Cout, b_mn=True, cs=_scalar(sc))
        # d what[n,k] = s_c(k) / (se_w[n] sc) sum_m A16[m,n] q[m,k]  +  sum_p (sum_b dY[b,p,n]) aft[p, k % KK]
        q16 = ops.codes_to_bf16(qcols, 1, M, K, K, 0, False, FMT)
        colscale = (se.repeat_interleave(KK) * sc[1]).contiguous()
        inv_sew = (1.0 / se_w).contiguous()
        dwhat = torch.zeros((Cout, K), dtype=torch.float32, device=dY.device)
        ops.gemm(GEMM_BWD, a16, (Cout, 0, 0, 0), q16, (K, 0, 0, 0), dwhat, (K, 0, 0), Cout, K, M, a_mn=True, b_mn=True, splits=0,
                 accumulate=True, rs=vec(inv_sew), cs=vec(colscale))
        dysum = dY.view(B, P, Cout).sum(0)                           # [P, Cout]
        dwhat.view(Cout, Cin, KK).add_((dysum.t() @ aftcols).unsqueeze(1))
        # d x_hat back in image layout, then the image quantizer's straight-through / scale / shift gradients
        dxhat = dxcols.view(B, gh, gw, Cin, kh, kw).permute(0, 3, 1, 4, 2, 5).contiguous().view(B * Cin, H * W)
        dx, ds, db4, daft = ops.lsq_bwd(dxhat, xc.view(B * Cin, H * W), b4, se, PER_ROW, Cin, 1, lo, hi, g)
        return ((dx.view_as(xc) if ctx.needs_input_grad[0] else None), db4, daft, ds, dwhat.view(Cout, Cin, kh, kw), None,
                dbias, None, None)


# ====================================================================================== attention core
def _pv_forward(qp, ldq, rowsum, qv, se_p, se_v, v_aft, B, N, H, C):
    """out[b,n,h*hd+j] = se_p[n] * (se_v[hj] * sum_d qp[z,n,d] qv[b,d,hj] + v_aft[hj] * sum_d qp[z,n,d])."""
    hd = C // H
    qvT = ops.codes_transpose(qv, B, N, C, C, N * C)                 # [B, C, NP16]
    npad = qvT.shape[-1]
    out = torch.empty((B, N, C), dtype=torch.float32, device=qv.device)
    ops.gemm(GEMM_I8, qp, (ldq, 0, N * ldq, H * N * ldq), qvT, (npad, 0, hd * npad, C * npad), out,
             (C, hd, N * C), N, hd, N, nb1=H, nb2=B,
             rs=vec(se_p, N), cs=vec(se_v, 0, hd), rt=vec(rowsum, 0, N, H * N), ct=vec(v_aft, 0, hd))
    return out


def _pv_backward_f16(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, qv16=None, qp16=None, sc=None, amax_dv=None,
                     amax_dp=None, skip_dp=False):
    """fp16 backward of P_hat V_hat with ONE copy A16[b,n,c] = fp16(dO * se_v[c] * se_p[n] * sc):
        dP_hat[z,n,d]  = 1/(se_p[n] sc) * sum_j A16[b,n,hj] qv[b,d,hj] + sum_j dO[b,n,hj] v_aft[hj]
        dv_hat[b,d,hj] = 1/(se_v[hj] sc) * sum_n qp[z,n,d] A16[b,n,hj]                    (both operands MN-major)"""
    hd = C // H
    if sc is None:
        sc = ops.absmax_scale(dO, B, N, C, C, N * C, cs=sv2[0], rs=sp2[0], rs_period=N, product=True)
    prep = ops.grad_prep(dO, B, N, C, C, N * C, cs=sv2[0], rs=sp2[0], rs_period=N, want_rm=True, u=v_aft, group=hd,
                         fmt=FMT, scale4=sc, rm_rowscale=True)
    a16 = prep["rm"]                                                     # [1, B, N, C]
    if qv16 is None:
        qv16 = ops.codes_to_bf16(qv, B, N, C, C, N * C, False, FMT)      # [B, N, C]
    dPq = None
    if not skip_dp:
        dPq = torch.empty((B * H, N, ldS), dtype=torch.float32, device=dO.device)
        ops.gemm(GEMM_BWD, a16, (C, 0, hd, N * C), qv16, (C, 0, hd, N * C), dPq, (ldS, N * ldS, H * N * ldS), N, N, hd,
                 nb1=H, nb2=B, rs=vec(sp2[1], N), cs=_scalar(sc), rt=vec(prep["rowdot"], 0, N, H * N), amax=amax_dp)
    if qp16 is None:
        qp16 = ops.codes_to_bf16(qp, B * H, N, ldq, ldq, N * ldq, False, FMT)    # [B*H, N, ldq]
    dvhat = torch.empty((B, N, C), dtype=torch.float32, device=dO.device)
    ops.gemm(GEMM_BWD, qp16, (ldq, 0, N * ldq, H * N * ldq), a16, (C, 0, hd, N * C), dvhat, (C, hd, N * C), N, hd, N,
             nb1=H, nb2=B, a_mn=True, b_mn=True, rs=_scalar(sc), cs=vec(sv2[1], 0, hd), amax=amax_dv)
    if skip_dp:       # the fused attention backward forms dP itself: hand it the operand, the rank-1 term and the range scale
        return (a16, prep["rowdot"], sc, qv16), dvhat
    return dPq, dvhat


def _pv_backward(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, qv16=None, qp16=None, sc=None, amax_dv=None, amax_dp=None):
    """Returns (dPq [B*H,N,ldS] fp32, dvhat [B,N,C] fp32). sp2 / sv2 = [scale, 1/scale] of the probability / V quantizer.
    qv16 / qp16: exact 16-bit copies of the codes left by the forward quantizer passes (fp16 mode)."""
    if F16:
        return _pv_backward_f16(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, qv16, qp16, sc, amax_dv, amax_dp)
    se_p, se_v = sp2[0], sv2[0]
    hd = C // H
    prep = ops.grad_prep(dO, B, N, C, C, N * C, cs=se_v, rs=se_p, rs_period=N, want_rm=True, want_t=True,
                         u=v_aft, group=hd, planes=PLANES)
    npad8 = prep["r_pad"]
    # dP_hat[z,n,d] = sum_j (dO[n,hj] se_v[hj]) qv[d,hj] + sum_j dO[n,hj] v_aft[hj]
    qv16 = ops.codes_to_bf16(qv, B, N, C, C, N * C, False)           # [B, N, C] bf16
    dPq = torch.empty((B * H, N, ldS), dtype=torch.float32, device=dO.device)
    ops.gemm(GEMM_BF16, prep["rm"], (C, B * N * C, hd, N * C), qv16, (C, 0, hd, N * C), dPq, (ldS, N * ldS, H * N * ldS),
             N, N, hd, k2=K2P, a_dual_delta=DD, nb1=H, nb2=B, rt=vec(prep["rowdot"], 0, N, H * N))
    # dv_hat[b,d,hj] = sum_n qp[z,n,d] (se_p[n] dO[b,n,hj])
    qpT = ops.codes_to_bf16(qp, B * H, N, ldq, ldq, N * ldq, True)   # [B*H, ldq, npad8]; rows d >= N are zero
    dvhat = torch.empty((B, N, C), dtype=torch.float32, device=dO.device)
    ops.gemm(GEMM_BF16, qpT, (npad8, 0, ldq * npad8, H * ldq * npad8), prep["t"],
             (npad8, B * C * npad8, hd * npad8, C * npad8), dvhat, (C, hd, N * C), N, hd, N, k2=PLANES, nb1=H, nb2=B)
    return dPq, dvhat


class QKRAttnCoreFn(_Fn):
    """QAttention_qkreparam.forward up to (not including) proj (attention.py:174-219), and the Swin variant
    (swin_attention_and_mlp.py:168-229) through `attn_bias` / `attn_mask`.  x is [B, N, C]."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, bv, x_b4, x_aft, s_x, v_b4, v_aft, s_v, k_b4, k_aft, s_k, s_p,
                attn_bias, attn_mask, H: int, wbits: int, abits: int, nW: int, link: Optional[MlpLink] = None):
        B, N, C = x.shape
        M = B * N
        hd = C // H
        dev = x.device
        lo, hi = levels(abits, False)
        _, hiu = levels(abits, True)
        scale = hd ** -0.5
        xc = x.contiguous()
        x2d = xc.view(M, C)
        # --- shared quantized input (LSQ_input, qlinear.py:21-26)
        g_x = grad_scale_factor(hi, B * C)
        sx2 = ops.lsq_effective_scale(s_x, g_x, recip=True)
        se_x = sx2[0]
        need_grad = _need_grad(ctx)
        f16 = FMT if (F16 and need_grad) else None     # every quantizer pass also leaves the exact fp16 copy the backward GEMMs read
        qx16 = qv16 = qk16 = qp16 = None
        qx = ops.lsq_quant(x2d, x_b4, se_x, PER_ROW, N, 1, lo, hi, fmt16=f16)
        if f16 is not None:
            qx, qx16 = qx
        # --- V branch (attention.py:179-186)
        wvc, cs_v, _, ct_v, _, ics_v, *wv16 = ops.statsq_codes(wv, wbits, aft=x_aft, bias=bv, want_inv=True, fmt16=f16)
        wv16 = wv16[0] if wv16 else None
        v_out = torch.empty((M, C), dtype=torch.float32, device=dev)
        ops.gemm(GEMM_I8, qx, (C, 0, 0, 0), wvc, (C, 0, 0, 0), v_out, (C, 0, 0), M, C, C,
                 rs=vec(se_x, N), cs=vec(cs_v), ct=vec(ct_v))
        g_v = grad_scale_factor(hi, B * N)
        sv2 = ops.lsq_effective_scale(s_v, g_v, recip=True)
        se_v = sv2[0]
        qv = ops.lsq_quant(v_out, v_b4, se_v, PER_COL, 1, 1, lo, hi, fmt16=f16)
        if f16 is not None:
            qv, qv16 = qv
        # --- QK branch: one StatsQ on the per-head product W_q^T W_k (attention.py:190-196)
        wqk = ops.wqk_compose(wq, wk, H)
        wqkc, cs_qk, _, ct_qk, _, ics_qk, *wqk16 = ops.statsq_codes(wqk, wbits, aft=x_aft, want_inv=True, fmt16=f16)
        wqk16 = wqk16[0] if wqk16 else None
        g_k = grad_scale_factor(hi, B * C)
        sk2 = ops.lsq_effective_scale(s_k, g_k, recip=True)          # [2, N*H], index n*H + h
        se_k = sk2[0]
        # The quantizer of qkx, and in the same pass the column term of the logits sum_c x_aft[c] qk[b,d,h,c]  ([M, H];
        # attention.py:210-213: S = x_hat . k_hat^T * scale; terms constant along the softmax axis are dropped, they cancel
        # exactly in softmax and in its gradient).
        fused_qkx = (FUSED_QKX and C % 32 == 0
                     and (not need_grad or (f16 is not None and FUSED16 and C % 128 == 0)))       # backward: the streaming fp16 pass
        qkx = qkx_res = None
        if fused_qkx:
            # ... as the EPILOGUE of the qkx GEMM: codes, their fp16 copy and the fp16 residual the backward needs leave the
            # kernel, the fp32 product x_hat W_qk^T (233 MB per DeiT-S block) is never written or re-read
            qk, qk16, qkx_res, ctS = ops.gemm_lsq(qx, wqkc, M, H * C, C, k_b4, sk2, N, H, lo, hi, rs=vec(se_x, N), cs=vec(cs_qk),
                                                  ct=vec(ct_qk), fmt16=f16, want_res=f16 is not None, dot_u=x_aft.repeat(H))
        else:
            qkx = torch.empty((M, H * C), dtype=torch.float32, device=dev)
            ops.gemm(GEMM_I8, qx, (C, 0, 0, 0), wqkc, (C, 0, 0, 0), qkx, (H * C, 0, 0), M, H * C, C,
                     rs=vec(se_x, N), cs=vec(cs_qk), ct=vec(ct_qk))
            r = ops.lsq_quant(qkx, k_b4, se_k, PER_ROW, N, H, lo, hi, fmt16=f16, dot_u=x_aft.repeat(H))   # codes [M, H*C]
            if f16 is not None:
                qk, qk16, ctS = r
            else:
                qk, ctS = r
        sk2_hn = sk2.view(2, N, H).transpose(1, 2).contiguous()      # [2, H, N]
        se_k_hn = sk2_hn[0]
        ldS = round_up(N, 4)
        g_p = grad_scale_factor(hiu, B * H * N)
        sp2 = ops.lsq_effective_scale(s_p, g_p, recip=True)
        se_p = sp2[0]
        fused_attn = (FUSED_ATTN and attn_bias is None and attn_mask is None and hd == 64 and N <= 208 and C <= 384
                      and (f16 is not None or not need_grad))
        rowstat = None
        if fused_attn:
            # --- scores, softmax, probability quantizer and P.V in one kernel (attention.py:210-219)
            qvT = ops.codes_transpose(qv, B, N, C, C, N * C)
            fused_bwd = need_grad and FUSED_ATTN_BWD and F16
            res = ops.qkr_attn_fwd(qx, qk, qvT, B, N, H, C, se_x, se_k, ctS, scale, se_p, hiu, se_v, v_aft,
                                   save_p=need_grad and not fused_bwd, fmt16=f16 if need_grad else None, want_rowstat=fused_bwd)
            out, qp, P, qp16 = res[:4]
            rowstat = res[5] if fused_bwd else None
            ldq = qp.shape[-1]
        else:
            cs_S = se_k_hn * scale
            ct_S = (ctS.view(B, N, H).permute(0, 2, 1) * cs_S.unsqueeze(0)).contiguous()   # [B, H, N]
            S = torch.empty((B * H, N, ldS), dtype=torch.float32, device=dev)
            ops.gemm(GEMM_I8, qx, (C, 0, 0, N * C), qk, (H * C, 0, C, N * H * C), S, (ldS, N * ldS, H * N * ldS),
                     N, N, C, nb1=H, nb2=B, rs=vec(se_x, N), cs=vec(cs_S, 0, N), ct=vec(ct_S, 0, N, H * N))
            # --- softmax + probability quantizer (attention.py:213-215)
            if f16 is not None and attn_bias is None and attn_mask is None:
                P, qp, rowsum, qp16 = ops.softmax_quant(S, N, H, se_p, hiu, save_p=need_grad, fmt16=f16)
            else:
                P, qp, rowsum = ops.softmax_quant(S, N, H, se_p, hiu, bias=attn_bias, mask=attn_mask, nW=nW, save_p=need_grad)
            ldq = qp.shape[-1]
            del S
            out = _pv_forward(qp, ldq, rowsum, qv, se_p, se_v, v_aft, B, N, H, C)
        if TAP is not None:
            P_tap = P
            if P_tap is None and fused_attn:      # debug only: the probabilities of the fused kernel (bit-identical re-run)
                P_tap = ops.qkr_attn_fwd(qx, qk, qvT, B, N, H, C, se_x, se_k, ctS, scale, se_p, hiu, se_v, v_aft, save_p=True)[2]
            qkx_tap = qkx
            if qkx_tap is None:                   # debug only: the product the quantizing epilogue never wrote
                qkx_tap = torch.empty((M, H * C), dtype=torch.float32, device=dev)
                ops.gemm(GEMM_I8, qx, (C, 0, 0, 0), wqkc, (C, 0, 0, 0), qkx_tap, (H * C, 0, 0), M, H * C, C,
                         rs=vec(se_x, N), cs=vec(cs_qk), ct=vec(ct_qk))
            TAP.append(("qkr", dict(x=x2d, x_b4=x_b4, se_x=se_x, qx=qx, wvc=wvc, v_out=v_out, v_b4=v_b4, se_v=se_v, qv=qv, wqkc=wqkc,
                                    qkx=qkx_tap, k_b4=k_b4, se_k=se_k, qk=qk, P=P_tap, se_p=se_p, qp=qp, out=out, B=B, N=N, H=H, C=C)))
        ctx.save_for_backward(xc, wq, wk, x_b4, x_aft, v_b4, v_aft, k_b4, k_aft, qx, sx2, wvc, cs_v, ics_v, v_out, qv, sv2,
                              wqkc, cs_qk, ics_qk, qkx if qkx is not None else qkx_res, qk, sk2, sk2_hn, P, qp, sp2, qx16, qv16, qk16, qp16, wv16, wqk16,
                              rowstat, ctS if rowstat is not None else None, wv)
        ctx.cfg = (B, N, C, H, lo, hi, hiu, scale, g_x, g_v, g_k, g_p, ldS, ldq, bv is not None,
                   attn_bias is not None, link)
        ctx.pro_gen = prologue.current_generation()
        if link is not None:
            link.cs, link.se, link.sc = se_v, se_p, None
        return out

    @staticmethod
    def backward(ctx, dO):
        prologue.check_generation(ctx.pro_gen)
        (xc, wq, wk, x_b4, x_aft, v_b4, v_aft, k_b4, k_aft, qx, sx2, wvc, cs_v, ics_v, v_out, qv, sv2, wqkc, cs_qk, ics_qk,
         qkx, qk, sk2, sk2_hn, P, qp, sp2, qx16, qv16, qk16, qp16, wv16, wqk16, rowstat, ctS, wv) = ctx.saved_tensors
        B, N, C, H, lo, hi, hiu, scale, g_x, g_v, g_k, g_p, ldS, ldq, has_bv, has_bias, link = ctx.cfg
        se_x, se_v, se_k, se_p, se_k_hn = sx2[0], sv2[0], sk2[0], sp2[0], sk2_hn[0]
        M = B * N
        dev = dO.device
        dO = dO.contiguous()
        # fp16 mode: the GEMMs that produce d v_hat / d k_hat track max |output| in their epilogue; the LSQ backward passes
        # of the V and qkx quantizers then write the fp16 operand of the next linear layer's backward GEMMs directly
        # (range scale from that bound), so d v_out / d qkx never exist in fp32 and ofq_grad_prep is not needed there
        fused16 = F16 and FUSED16 and C % 128 == 0      # streaming layout of the (token, head)-segmented qkx pass
        amax = ops.scratch_zeros(3, dev) if F16 else None      # max |d v_hat|, |d k_hat|, |dP_hat|
        fused_bwd = rowstat is not None
        if fused_bwd:
            (a16_o, rowdot, sc_o, qv16), dvhat = _pv_backward_f16(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, qv16, qp16,
                                                                 link.sc if link is not None else None,
                                                                 amax_dv=amax[0:1] if fused16 else None, skip_dp=True)
        else:
            dPq, dvhat = _pv_backward(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, qv16, qp16,
                                      link.sc if link is not None else None, amax_dv=amax[0:1] if fused16 else None,
                                      amax_dp=amax[2:3] if F16 else None)
        # --- V quantizer and V linear
        dxhat = torch.empty((M, C), dtype=torch.float32, device=dev)
        if fused16:
            sc_v = ops.scale_from_max(amax[0:1], v1=cs_v, v2=se_x, product=True)
            _, ds_v, dvb4, dvaft, a16_v = ops.lsq_bwd(dvhat.view(M, C), v_out, v_b4, se_v, PER_COL, 1, 1, lo, hi, g_v,
                                                      out16=(FMT, cs_v, se_x, N, sc_v), want_dx=False)
            dWv, dbv, qx_op = _linear_backward_f16(None, qx, wvc, (cs_v, ics_v), sx2, N, x_aft, dxhat, False, qx16, sc=sc_v,
                                                   a16=a16_v, colsum=dvb4, wc16=wv16, dw_for=wv)
        else:
            dv_out, ds_v, dvb4, dvaft, *sc_v = ops.lsq_bwd(dvhat.view(M, C), v_out, v_b4, se_v, PER_COL, 1, 1, lo, hi, g_v,
                                                           next_scale=(cs_v, se_x, 1.0, True) if F16 else None)
            dWv, dbv, qx_op = _linear_backward(dv_out, qx, wvc, (cs_v, ics_v), sx2, N, x_aft, dxhat, False, qx16,
                                               sc=sc_v[0] if sc_v else None, dw_for=wv, **({"wc16": wv16} if F16 else {}))
        # --- softmax + probability quantizer, then the two score GEMMs
        if F16:
            if fused_bwd:
                # logits recomputed from the codes, dP = dO v^T, softmax / quantizer backward: one kernel, nothing of it in HBM
                dS16, ldo, colsum_dS, ds_p, sc = ops.qkr_attn_bwd(qx, qk, a16_o, qv16, FMT, B, N, H, C, se_x, se_k, ctS, scale, sp2, hiu,
                                                                rowstat, rowdot, sc_o, se_v, v_aft, -lo, g_p)
                dS32 = None
            else:
                # |dS| = |alpha P (dP - sum P dP)| <= 2 alpha max|dPq|; ONE copy dS16[b,h,n,d] = fp16(dS se_k[h,d] se_x[n] sc)
                # (max |dP_hat| comes from the epilogue of the GEMM that produced it: no pass over dP_hat)
                sc = ops.scale_from_max(amax[2:3], v1=se_k_hn, v2=se_x, mult=2.0 * scale, product=True)
                dS16, _, ldo, colsum_dS, ds_p, dS32 = ops.softmax_quant_bwd(dPq, P, N, H, se_p, hiu, scale, g_p, se_k_hn, True,
                                                                          se_x, want_ds32=has_bias, fmt=FMT, scale4=sc,
                                                                          single=True)
                del dPq
            slab = N * ldo
            # d x_hat[b,n,c] += 1/(se_x[n] sc) sum_h sum_d dS16[b,h,n,d] qk[b,d,h,c]      (heads = outer-K, B MN-major)
            if qk16 is None:
                qk16 = ops.codes_to_bf16(qk, 1, M, H * C, H * C, 0, False, FMT)     # [1, M, H*C] = [b][d][h][c]
            ops.gemm(GEMM_BWD, dS16, (ldo, slab, H * slab, 0), qk16, (H * C, C, N * H * C, 0), dxhat, (C, N * C, 0),
                     N, C, N, k2=H, nb1=B, accumulate=True, b_mn=True, rs=vec(sx2[1], N), cs=_scalar(sc))
            # d k_hat[b,d,h,c] = 1/(se_k[h,d] sc) sum_n dS16[b,h,n,d] qx[b,n,c] + colsum_dS[z,d] x_aft[c]
            dkhat = torch.empty((M, H * C), dtype=torch.float32, device=dev)
            ops.gemm(GEMM_BWD, dS16, (ldo, 0, slab, H * slab), qx_op, (C, 0, 0, N * C), dkhat, (H * C, C, N * H * C),
                     N, C, N, nb1=H, nb2=B, a_mn=True, b_mn=True, rs=vec(sk2_hn[1], 0, N), cs=_scalar(sc),
                     rt=vec(colsum_dS, 0, N, H * N), ct=vec(x_aft), amax=amax[1:2] if fused16 else None)
            del dS16
        else:
            dSa, dSbT, ldo, colsum_dS, ds_p, dS32 = ops.softmax_quant_bwd(dPq, P, N, H, se_p, hiu, scale, g_p, se_k_hn, True,
                                                                        se_x, want_ds32=has_bias, planes=PLANES)
            slab = N * ldo                      # one (b, plane, h) slab of dSa / dSbT, laid out [B, PLANES, H, N, ldo]
            del dPq
            # --- scores: d x_hat += sum_h dS (k_hat)   (outer-K loop over heads)
            qkT = ops.codes_to_bf16(qk, B, N, H * C, H * C, N * H * C, True)      # [B, H*C, npad8]
            npad8 = qkT.shape[-1]
            if DUAL:     # planes of head h are outer-K slices h and h + H of dSa [B, 2, H, N, ldo]
                ops.gemm(GEMM_BF16, dSa, (ldo, slab, PLANES * H * slab, 0), qkT, (npad8, C * npad8, H * C * npad8, 0),
                         dxhat, (C, N * C, 0), N, C, N, k2=H, a_dual_delta=H, nb1=B, accumulate=True)
            else:
                ops.gemm(GEMM_BF16, dSa, (ldo, slab, PLANES * H * slab, 0), qkT, (npad8, C * npad8, H * C * npad8, 0),
                         dxhat, (C, N * C, 0), N, C, N, k2=PLANES * H, nb1=B, accumulate=True, b_k2mod=H)
            # --- scores: d k_hat[b,d,h,c] = sum_n dS[n,d] x_hat[n,c]
            qxT_b = ops.codes_to_bf16(qx, B, N, C, C, N * C, True)                # [B, C, npad8]
            dkhat = torch.empty((M, H * C), dtype=torch.float32, device=dev)
            ops.gemm(GEMM_BF16, dSbT, (ldo, H * slab, slab, PLANES * H * slab), qxT_b, (npad8, 0, 0, C * npad8), dkhat,
                     (H * C, C, N * H * C), N, C, N, k2=K2P, a_dual_delta=DD, nb1=H, nb2=B, rt=vec(colsum_dS, 0, N, H * N),
                     ct=vec(x_aft))
            del dSa, dSbT
        # --- qkx quantizer and the qkx "linear" layer (weight = StatsQ(W_q^T W_k), no bias)
        # d(move_qkx_aft) is analytically zero: the shift adds a term to the logits that is constant along the softmax axis
        dkaft = torch.zeros_like(k_aft)
        if fused16:
            sc_k = ops.scale_from_max(amax[1:2], v1=cs_qk, v2=se_x, product=True)
            # (qkx is the fp16 residual plane when the forward quantized in the GEMM epilogue)
            _, ds_k, dkb4, _, a16_k = ops.lsq_bwd(dkhat, qkx, k_b4, se_k, PER_ROW, N, H, lo, hi, g_k, want_aft=False, zero_sum=True,
                                                  out16=(FMT, cs_qk, se_x, N, sc_k), want_dx=False,
                                                  act=ops.ACT_RES16 if qkx.dtype == torch.float16 else ops.ACT_NONE)
            del dkhat
            # colsum(d qkx) = sum over rows of the masked gradient = d(move_qkx_b4) (its zero-sum form differs from the plain
            # sum by sum_rows d k_hat, which is analytically zero)
            dWqk, _, _ = _linear_backward_f16(None, qx, wqkc, (cs_qk, ics_qk), sx2, N, x_aft, dxhat, True, qx_op, sc=sc_k,
                                              a16=a16_k, colsum=dkb4, wc16=wqk16)
        else:
            dqkx, ds_k, dkb4, _, *sc_k = ops.lsq_bwd(dkhat, qkx, k_b4, se_k, PER_ROW, N, H, lo, hi, g_k, want_aft=False,
                                                     zero_sum=True, next_scale=(cs_qk, se_x, 1.0, True) if F16 else None)
            del dkhat
            dWqk, _, _ = _linear_backward(dqkx, qx, wqkc, (cs_qk, ics_qk), sx2, N, x_aft, dxhat, True, qx_op,
                                          sc=sc_k[0] if sc_k else None, **({"wc16": wqk16} if F16 else {}))
        dwq, dwk = torch.empty_like(wq), torch.empty_like(wk)
        with ops.side_stream(dWqk, wq, wk, dwq, dwk):      # ordered after the dW_qk GEMM on the side stream
            ops.wqk_compose_bwd(dWqk, wq, wk, H, out=(dwq, dwk))
        # --- shared input quantizer
        dx, ds_x, dxb4, dxaft = ops.lsq_bwd(dxhat, xc.view(M, C), x_b4, se_x, PER_ROW, N, 1, lo, hi, g_x)
        dbias = None
        if has_bias:
            dbias = dS32[..., :N].reshape(B, H, N, N).sum(0)
        ops.side_join()
        return (dx.view(B, N, C), dwq, dwk, dWv, (dbv if has_bv else None), dxb4, dxaft, ds_x, dvb4, dvaft, ds_v,
                dkb4, dkaft, ds_k, ds_p, dbias, None, None, None, None, None, None)


class QAttnCoreFn(_Fn):
    """QAttention.forward between the qkv QLinear and proj (attention.py:70-102) / QAttention_swin
    (swin_attention_and_mlp.py:172-229).  qkv is [B, N, 3C] laid out (3, H, hd) along the last dim."""

    @staticmethod
    def forward(ctx, qkv, b4, s_q, s_k, s_v, q_aft, k_aft, v_aft, s_p, attn_bias, attn_mask, H: int, abits: int, nW: int,
                link: Optional[MlpLink] = None):
        B, N, C3 = qkv.shape
        C = C3 // 3
        M = B * N
        hd = C // H
        dev = qkv.device
        lo, hi = levels(abits, False)
        _, hiu = levels(abits, True)
        scale = hd ** -0.5
        qkvc = qkv.contiguous()
        q2d = qkvc.view(M, C3)
        g_qk = grad_scale_factor(hi, B * C)           # 4-D (B,H,N,hd): B*H*hd elements per token scale
        sq2 = ops.lsq_effective_scale(s_q, g_qk, recip=True)
        sk2 = ops.lsq_effective_scale(s_k, g_qk, recip=True)
        g_v = grad_scale_factor(hi, B * N)
        sv2 = ops.lsq_effective_scale(s_v, g_v, recip=True)
        se_q, se_k, se_v = sq2[0], sk2[0], sv2[0]
        qq = ops.lsq_quant(q2d[:, 0:C], b4[0:C], se_q, PER_ROW, N, 1, lo, hi)
        qk = ops.lsq_quant(q2d[:, C:2 * C], b4[C:2 * C], se_k, PER_ROW, N, 1, lo, hi)
        qv = ops.lsq_quant(q2d[:, 2 * C:], b4[2 * C:], se_v, PER_COL, 1, 1, lo, hi)
        # S = q_hat k_hat^T * scale with q_hat = qq se_q[n] + q_aft, k_hat = qk se_k[d] + k_aft (row-only terms dropped)
        ctS = ops.codes_rowdot(qk, H, q_aft)                          # [M, H]: sum_j q_aft[hj] qk[b,d,hj]
        cs_S = se_k * scale                                           # [N]
        ct_S = (ctS.view(B, N, H).permute(0, 2, 1) * cs_S.view(1, 1, N)).contiguous()   # [B, H, N]
        ldS = round_up(N, 4)
        S = torch.empty((B * H, N, ldS), dtype=torch.float32, device=dev)
        ops.gemm(GEMM_I8, qq, (C, 0, hd, N * C), qk, (C, 0, hd, N * C), S, (ldS, N * ldS, H * N * ldS),
                 N, N, hd, nb1=H, nb2=B, rs=vec(se_q, N), cs=vec(cs_S), ct=vec(ct_S, 0, N, H * N))
        g_p = grad_scale_factor(hiu, B * H * N)
        sp2 = ops.lsq_effective_scale(s_p, g_p, recip=True)
        se_p = sp2[0]
        need_grad = _need_grad(ctx)
        P, qp, rowsum = ops.softmax_quant(S, N, H, se_p, hiu, bias=attn_bias, mask=attn_mask, nW=nW, save_p=need_grad)
        ldq = qp.shape[-1]
        del S
        out = _pv_forward(qp, ldq, rowsum, qv, se_p, se_v, v_aft, B, N, H, C)
        if TAP is not None:
            TAP.append(("attn", dict(qkv=q2d, b4=b4, se_q=se_q, se_k=se_k, se_v=se_v, qq=qq, qk=qk, qv=qv, P=P, se_p=se_p, qp=qp,
                                     out=out, B=B, N=N, H=H, C=C)))
        ctx.save_for_backward(qkvc, b4, q_aft, k_aft, v_aft, qq, qk, qv, sq2, sk2, sv2, P, qp, sp2)
        ctx.cfg = (B, N, C, H, lo, hi, hiu, scale, g_qk, g_v, g_p, ldS, ldq, attn_bias is not None, link)
        if link is not None:
            link.cs, link.se, link.sc = se_v, se_p, None
        return out

    @staticmethod
    def backward(ctx, dO):
        qkvc, b4, q_aft, k_aft, v_aft, qq, qk, qv, sq2, sk2, sv2, P, qp, sp2 = ctx.saved_tensors
        B, N, C, H, lo, hi, hiu, scale, g_qk, g_v, g_p, ldS, ldq, has_bias, link = ctx.cfg
        se_q, se_k, se_v, se_p = sq2[0], sk2[0], sv2[0], sp2[0]
        M = B * N
        hd = C // H
        dev = dO.device
        dO = dO.contiguous()
        q2d = qkvc.view(M, 3 * C)
        dPq, dvhat = _pv_backward(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, sc=link.sc if link is not None else None)
        dqhat = torch.empty((M, C), dtype=torch.float32, device=dev)
        dkhat = torch.empty((M, C), dtype=torch.float32, device=dev)
        if F16:
            sc = ops.absmax_scale(dPq, B * H, N, N, ldS, N * ldS, v1=se_k, v2=se_q, mult=2.0 * scale, product=True)
            dS16, _, ldo, colsum_dS, ds_p, dS32 = ops.softmax_quant_bwd(dPq, P, N, H, se_p, hiu, scale, g_p, se_k, False,
                                                                      se_q, want_ds32=has_bias, fmt=FMT, scale4=sc,
                                                                      single=True)
            slab = N * ldo
            del dPq
            # dq_hat[b,n,hj] = 1/(se_q[n] sc) sum_d dS16[b,h,n,d] qk[b,d,hj]                          (B MN-major)
            qk16 = ops.codes_to_bf16(qk, 1, M, C, C, 0, False, FMT)                 # [1, M, C] = [b][d][hj]
            ops.gemm(GEMM_BWD, dS16, (ldo, 0, slab, H * slab), qk16, (C, 0, hd, N * C), dqhat, (C, hd, N * C), N, hd, N,
                     nb1=H, nb2=B, b_mn=True, rs=vec(sq2[1], N), cs=_scalar(sc))
            # dk_hat[b,d,hj] = 1/(se_k[d] sc) sum_n dS16[b,h,n,d] qq[b,n,hj] + colsum_dS[z,d] q_aft[hj]   (both MN-major)
            qq16 = ops.codes_to_bf16(qq, 1, M, C, C, 0, False, FMT)
            # the rank-1 term is not scaled by rs / cs, so the scalar rides on cs and 1/se_k on rs
            ops.gemm(GEMM_BWD, dS16, (ldo, 0, slab, H * slab), qq16, (C, 0, hd, N * C), dkhat, (C, hd, N * C), N, hd, N,
                     nb1=H, nb2=B, a_mn=True, b_mn=True, rs=vec(sk2[1], N), cs=_scalar(sc),
                     rt=vec(colsum_dS, 0, N, H * N), ct=vec(q_aft, 0, hd))
            del dS16
        else:
            dSa, dSbT, ldo, colsum_dS, ds_p, dS32 = ops.softmax_quant_bwd(dPq, P, N, H, se_p, hiu, scale, g_p, se_k, False,
                                                                        se_q, want_ds32=has_bias, planes=PLANES)
            slab = N * ldo
            del dPq
            # dq_hat[b,n,hj] = sum_d (dS se_k[d]) qk[b,d,hj]
            qkT = ops.codes_to_bf16(qk, B, N, C, C, N * C, True)          # [B, C, npad8]
            npad8 = qkT.shape[-1]
            ops.gemm(GEMM_BF16, dSa, (ldo, H * slab, slab, PLANES * H * slab), qkT, (npad8, 0, hd * npad8, C * npad8), dqhat,
                     (C, hd, N * C), N, hd, N, k2=K2P, a_dual_delta=DD, nb1=H, nb2=B)
            # dk_hat[b,d,hj] = sum_n (dS se_q[n]) qq[b,n,hj] + colsum_dS[z,d] q_aft[hj]
            qqT = ops.codes_to_bf16(qq, B, N, C, C, N * C, True)
            ops.gemm(GEMM_BF16, dSbT, (ldo, H * slab, slab, PLANES * H * slab), qqT, (npad8, 0, hd * npad8, C * npad8), dkhat,
                     (C, hd, N * C), N, hd, N, k2=K2P, a_dual_delta=DD, nb1=H, nb2=B, rt=vec(colsum_dS, 0, N, H * N),
                     ct=vec(q_aft, 0, hd))
            del dSa, dSbT
        dq, ds_q, db4_q, daft_q = ops.lsq_bwd(dqhat, q2d[:, 0:C], b4[0:C], se_q, PER_ROW, N, 1, lo, hi, g_qk)
        # d(move_k_aft) is analytically zero (its logit term q_hat . k_aft is constant along the softmax axis)
        dk, ds_k, db4_k, _ = ops.lsq_bwd(dkhat, q2d[:, C:2 * C], b4[C:2 * C], se_k, PER_ROW, N, 1, lo, hi, g_qk, want_aft=False,
                                         zero_sum=True)
        daft_k = torch.zeros_like(k_aft)
        dv, ds_v, db4_v, daft_v = ops.lsq_bwd(dvhat.view(M, C), q2d[:, 2 * C:], b4[2 * C:], se_v, PER_COL, 1, 1, lo, hi, g_v)
        dqkv = torch.cat((dq, dk, dv), dim=1).view(B, N, 3 * C)
        db4 = torch.cat((db4_q, db4_k, db4_v))
        dbias = dS32[..., :N].reshape(B, H, N, N).sum(0) if has_bias else None
        return dqkv, db4, ds_q, ds_k, ds_v, daft_q, daft_k, daft_v, ds_p, dbias, None, None, None, None, None
