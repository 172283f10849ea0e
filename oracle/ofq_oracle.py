"""CPU oracle for the OFQ quantization-aware-training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (plain torch fp32 on CPU, autograd for the gradients) of the reference's
fake-quant algorithm.  It is the checker: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The product path (ofq_b200/) never does.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
outputs of the reference itself, generated in the build container by tests/golden/make_golden.py (which
imports /root/reference under import stubs) and committed as tests/golden/*.npz; tests/test_oracle_golden.py
replays them.  Everything is written in functional style over a flat dict of tensors that uses the
reference's state-dict key names, so a reference checkpoint drops straight in.

Citations are file:line in nbasyl/OFQ.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

EPS_FLOOR = 1e-5  # lsq.py:593 `clip(alpha, 1e-5)`

# Test instrumentation: when a dict, every LSQ quantizer call made through the layer functions below records
# {"x": its input, "se": the effective step size, "codes": the integer codes} under the reference's module name
# (e.g. "blocks.0.attn.quan_a_qkx_fn"), and the layers record a few named intermediates ("<module>.@v_out", ...).
TAPS: Optional[dict] = None


# Matrix products: False = exactly the reference's ops (torch fp32 sgemm / bmm / einsum: bit-identical to the reference on
# the same machine). True = the same products accumulated in float64 and rounded to fp32 ONCE, i.e. the correctly rounded
# value of the reference's arithmetic. The reference's sgemm carries ~1e-6 of accumulated rounding error (K = 384..1536
# terms), enough to flip a handful of 2-bit codes that sit on a rounding tie; tests use the exact mode to separate those
# ties (a property of the reference's BLAS) from the CUDA path, whose integer GEMMs are exact (tests/test_oracle_fullsize.py).
EXACT_GEMM = False


def _linear(x: Tensor, w: Tensor) -> Tensor:
    if EXACT_GEMM:
        return F.linear(x.double(), w.double()).float()
    return F.linear(x, w)


def _matmul(a: Tensor, b: Tensor) -> Tensor:
    if EXACT_GEMM:
        return (a.double() @ b.double()).float()
    return a @ b


def _einsum(eq: str, a: Tensor, b: Tensor) -> Tensor:
    if EXACT_GEMM:
        return torch.einsum(eq, a.double(), b.double()).float()
    return torch.einsum(eq, a, b)


def _tap(name: Optional[str], **kw) -> None:
    if TAPS is not None and name is not None:
        TAPS[name] = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}


# ------------------------------------------------------------------------------------------ STE helpers
def ste_round(x: Tensor) -> Tensor:
    """lsq.py:11-14 round_pass: value round(x) (half-to-even), gradient identity."""
    r = x.round()
    return (r - x).detach() + x


def grad_scaled(x: Tensor, g: float) -> Tensor:
    """lsq.py:6-9 grad_scale: value (x - x*g) + x*g (NOT always bit-equal to x), gradient g."""
    xg = x * g
    return (x - xg).detach() + xg


def floor_clip(x: Tensor, eps: float = EPS_FLOOR) -> Tensor:
    """lsq.py:16-18 clip: value max-like where(x > eps, x, eps), gradient identity."""
    floor = torch.tensor(eps, dtype=torch.float32, device=x.device)
    xc = torch.where(x > floor, x, floor)
    return x - x.detach() + xc.detach()


# ------------------------------------------------------------------------------------------ StatsQ
def statsq_scale(w: Tensor) -> Tensor:
    """statsq.py:137-142: per-output-channel scale 2*mean|w| (2-D weight), detached."""
    assert w.dim() == 2
    return (2 * torch.mean(w.abs(), dim=1, keepdim=True)).detach()


def statsq_pre_round(w: Tensor, bits: int) -> Tuple[Tensor, Tensor]:
    """statsq.py:144-147 (also cga.py:454-459): returns (b4_round, scale). clip_val is the frozen 2.0."""
    sf = statsq_scale(w)
    clip_val = torch.tensor([2.0], dtype=torch.float32, device=w.device)
    scaled = w / sf
    clipped = torch.clamp(scaled, min=(-clip_val / 2), max=(clip_val / 2) - 1e-6)
    n = float(2 ** (bits - 1))
    return clipped * n - 0.5, sf


def statsq_codes(w: Tensor, bits: int) -> Tuple[Tensor, Tensor]:
    """Integer view of StatsQ: odd codes 2k+1 in [-(2^b-1), 2^b-1] and the scale sf (rows,1)."""
    b4, sf = statsq_pre_round(w, bits)
    k = torch.round(b4)
    return (2 * k + 1).to(torch.int32), sf


def statsq(w: Tensor, bits: int) -> Tensor:
    """StatsQuantizer.forward, statsq.py:133-150.  Gradient: identity for every element (line 148)."""
    b4, sf = statsq_pre_round(w, bits)
    n = float(2 ** (bits - 1))
    wq = sf * ((torch.round(b4) + 0.5) / n)
    return wq.detach() - w.detach() + w


# ------------------------------------------------------------------------------------------ LSQ
def lsq_levels(bit: int, all_positive: bool) -> Tuple[int, int]:
    """lsq.py:519-534."""
    if all_positive:
        return (0, 1) if bit == 1 else (0, 2 ** bit - 1)
    return (-1, 1) if bit == 1 else (-(2 ** (bit - 1)), 2 ** (bit - 1) - 1)


def lsq_init_rows(x: Tensor, hi: int, all_positive: bool) -> Tensor:
    """LsqQuantizer.init_from, lsq.py:544-569: one scale per index of dim -2."""
    f = 4 if all_positive else 2
    a = x.detach().abs().mean(dim=-1)
    if x.dim() == 3:
        a = a.mean(dim=0)
    elif x.dim() == 4:
        a = a.mean(dim=0).mean(dim=0)
    else:
        assert x.dim() == 2
    return f * a / (hi ** 0.5)


def lsq_init_cols(x: Tensor, hi: int, all_positive: bool) -> Tensor:
    """LsqQuantizer4v.init_from, lsq.py:730-754: one scale per last-dim channel."""
    f = 4 if all_positive else 2
    a = x.detach().abs()
    for _ in range(x.dim() - 1):
        a = a.mean(dim=0)
    return f * a / (hi ** 0.5)


def _lsq_core(x: Tensor, alpha: Tensor, g: float, lo: int, hi: int, bit: int, all_positive: bool,
              tap: Optional[str] = None) -> Tensor:
    s = grad_scaled(floor_clip(alpha), g)
    v = x / s
    if bit == 1 and not all_positive:
        v = torch.sign(v)
    else:
        v = ste_round(torch.clamp(v, lo, hi))
    _tap(tap, x=x, se=s, codes=v.detach().to(torch.int8), lo=lo, hi=hi)
    return v * s


def lsq_rows(x: Tensor, s: Tensor, bit: int, all_positive: bool, tap: Optional[str] = None) -> Tensor:
    """LsqQuantizer.forward (per_channel=True), lsq.py:571-602: scale indexed by dim -2."""
    lo, hi = lsq_levels(bit, all_positive)
    if x.dim() == 3:
        g = 1.0 / ((hi * x.shape[0] * x.shape[-1]) ** 0.5)
    elif x.dim() == 2:
        g = 1.0 / ((hi * x.shape[-1]) ** 0.5)
    else:
        assert x.dim() == 4
        g = 1.0 / ((hi * x.shape[0] * x.shape[1] * x.shape[-1]) ** 0.5)
    return _lsq_core(x, s.unsqueeze(-1), g, lo, hi, bit, all_positive, tap)


def lsq_cols(x: Tensor, s: Tensor, bit: int, all_positive: bool, tap: Optional[str] = None) -> Tensor:
    """LsqQuantizer4v.forward, lsq.py:757-790: scale indexed by the last dim."""
    lo, hi = lsq_levels(bit, all_positive)
    if x.dim() == 3:
        g = 1.0 / ((hi * x.shape[0] * x.shape[1]) ** 0.5)
        a = s.unsqueeze(0).unsqueeze(1)
    else:
        assert x.dim() == 4
        g = 1.0 / ((hi * x.shape[0] * x.shape[1] * x.shape[2]) ** 0.5)
        a = s.unsqueeze(0).unsqueeze(1).unsqueeze(2)
    return _lsq_core(x, a, g, lo, hi, bit, all_positive, tap)


def lsq_codes_rows(x: Tensor, s: Tensor, bit: int, all_positive: bool) -> Tensor:
    """Integer codes round(clamp(x/s', lo, hi)) the forward of lsq_rows uses (for bit-exact code parity)."""
    lo, hi = lsq_levels(bit, all_positive)
    with torch.no_grad():
        q = lsq_rows(x, s, bit, all_positive)
        if x.dim() == 3:
            g = 1.0 / ((hi * x.shape[0] * x.shape[-1]) ** 0.5)
        elif x.dim() == 2:
            g = 1.0 / ((hi * x.shape[-1]) ** 0.5)
        else:
            g = 1.0 / ((hi * x.shape[0] * x.shape[1] * x.shape[-1]) ** 0.5)
        se = grad_scaled(floor_clip(s.unsqueeze(-1)), g)
        return torch.round(torch.clamp(x / se, lo, hi)).to(torch.int32)


def _get_scale(P: Params, key: str, init_fn) -> Tensor:
    """Lazy data-dependent creation of `s` on first use (lsq.py:573-574): stored back into the dict."""
    if key not in P or P[key] is None:
        P[key] = init_fn().clone().requires_grad_(True)
    return P[key]


# ------------------------------------------------------------------------------------------ layers
def lsq_input(x: Tensor, P: Params, pre: str, bit: int, all_positive: bool = False) -> Tensor:
    """LSQ_input / the input side of QLinear: move_b4 -> LsqQuantizer -> move_aft (qlinear.py:21-26, 66-68)."""
    _, hi = lsq_levels(bit, all_positive)
    x = x + P[pre + "move_b4.bias"].expand_as(x)
    s = _get_scale(P, pre + "input_quant_fn.s", lambda: lsq_init_rows(x, hi, all_positive))
    x = lsq_rows(x, s, bit, all_positive, tap=pre + "input_quant_fn")
    return x + P[pre + "move_aft.bias"].expand_as(x)


def qlinear(x: Tensor, P: Params, pre: str, wbits: int, abits: int, symmetric: bool = True) -> Tensor:
    """QLinear.forward, qlinear.py:58-73."""
    w = statsq(P[pre + "weight"], wbits)
    xq = lsq_input(x, P, pre, abits, all_positive=not symmetric)
    out = _linear(xq, w)
    out = out + P[pre + "bias"].view(1, -1).expand_as(out)
    _tap(pre + "@out", out=out)
    return out


def qmlp(x: Tensor, P: Params, pre: str, wbits: int, abits: int) -> Tensor:
    """QMLP.forward / QMLP_swin.forward, qlinear.py:123-136: fc1 signed input, GELU, fc2 unsigned input."""
    x = qlinear(x, P, pre + "fc1.", wbits, abits, symmetric=True)
    x = F.gelu(x)
    return qlinear(x, P, pre + "fc2.", wbits, abits, symmetric=False)


def _softmax_quant(attn: Tensor, P: Params, pre: str, abits: int) -> Tensor:
    _, hi = lsq_levels(abits, True)
    prob = F.softmax(attn, dim=-1)
    s = _get_scale(P, pre + "quan_a_softmax_fn.s", lambda: lsq_init_rows(prob, hi, True))
    return lsq_rows(prob, s, abits, True, tap=pre + "quan_a_softmax_fn")


def qattention(x: Tensor, P: Params, pre: str, heads: int, wbits: int, abits: int,
               bias: Optional[Tensor] = None) -> Tensor:
    """QAttention.forward (plain quantized q/k/v), attention.py:67-105. `bias` is an additive pre-softmax
    term (Swin relative-position bias + shift mask, swin_attention_and_mlp.py:201-223)."""
    B, N, C = x.shape
    hd = C // heads
    _, hi = lsq_levels(abits, False)
    qkv = qlinear(x, P, pre + "qkv.", wbits, abits, symmetric=True)
    qkv = qkv + P[pre + "move_qkv_b4.bias"].expand_as(qkv)
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    sq = _get_scale(P, pre + "quan_a_q_fn.s", lambda: lsq_init_rows(q, hi, False))
    q = lsq_rows(q, sq, abits, False, tap=pre + "quan_a_q_fn")
    sk = _get_scale(P, pre + "quan_a_k_fn.s", lambda: lsq_init_rows(k, hi, False))
    k = lsq_rows(k, sk, abits, False, tap=pre + "quan_a_k_fn")
    v = v.permute(0, 2, 1, 3).reshape(B, N, C)
    sv = _get_scale(P, pre + "quan_a_v_fn.s", lambda: lsq_init_cols(v, hi, False))
    v = lsq_cols(v, sv, abits, False, tap=pre + "quan_a_v_fn")
    q = q.permute(0, 2, 1, 3).reshape(B, N, C) + P[pre + "move_q_aft.bias"]
    k = k.permute(0, 2, 1, 3).reshape(B, N, C) + P[pre + "move_k_aft.bias"]
    v = v + P[pre + "move_v_aft.bias"]
    q = q.reshape(B, N, heads, hd).permute(0, 2, 1, 3)
    k = k.reshape(B, N, heads, hd).permute(0, 2, 1, 3)
    v = v.reshape(B, N, heads, hd).permute(0, 2, 1, 3)
    attn = _matmul(q, k.transpose(-2, -1).contiguous()) * (hd ** -0.5)
    if bias is not None:
        attn = attn + bias
    prob = _softmax_quant(attn, P, pre, abits)
    out = _matmul(prob, v).transpose(1, 2).reshape(B, N, C)
    _tap(pre + "@core_out", out=out)
    return qlinear(out, P, pre + "proj.", wbits, abits, symmetric=True)


def wqk_compose(wq: Tensor, wk: Tensor, heads: int) -> Tensor:
    """attention.py:190-194: per-head W_q[h]^T @ W_k[h], stacked to (heads*C, C)."""
    C = wq.shape[1]
    mq = wq.reshape(heads, wq.shape[0] // heads, C)
    mk = wk.reshape(heads, wk.shape[0] // heads, C)
    return _matmul(mq.transpose(-2, -1).contiguous(), mk).reshape(heads * C, C)


def qattention_qkr(x: Tensor, P: Params, pre: str, heads: int, wbits: int, abits: int,
                   bias: Optional[Tensor] = None) -> Tensor:
    """QAttention_qkreparam.forward, attention.py:174-222 (== _4_cga variant in value and gradient,
    SURVEY.md §8a row 5)."""
    B, N, C = x.shape
    hd = C // heads
    _, hi = lsq_levels(abits, False)
    xq = lsq_input(x, P, pre + "quant_x_4_qkv.", abits, all_positive=False)
    # V branch
    wv = statsq(P[pre + "v.weight"], wbits)
    v = _linear(xq, wv)
    v = v + P[pre + "v.bias"].view(1, -1).expand_as(v)
    v = v + P[pre + "move_v_b4.bias"]
    sv = _get_scale(P, pre + "quan_a_v_fn.s", lambda: lsq_init_cols(v, hi, False))
    v = lsq_cols(v, sv, abits, False, tap=pre + "quan_a_v_fn")
    v = v + P[pre + "move_v_aft.bias"]
    v = v.reshape(B, N, heads, hd).permute(0, 2, 1, 3)
    # QK branch: one StatsQ on the per-head product
    wqk = statsq(wqk_compose(P[pre + "q.weight"], P[pre + "k.weight"], heads), wbits).reshape(heads, C, C)
    qkx = _einsum("HDC, BCN -> BHDN", wqk, xq.transpose(-2, -1).contiguous())
    qkx = qkx.permute(0, 3, 1, 2).reshape(B, N, heads * C)
    qkx = qkx + P[pre + "move_qkx_b4.bias"]
    qkx = qkx.reshape(B, N * heads, C)
    sk = _get_scale(P, pre + "quan_a_qkx_fn.s", lambda: lsq_init_rows(qkx, hi, False))
    qkx = lsq_rows(qkx, sk, abits, False, tap=pre + "quan_a_qkx_fn")
    qkx = qkx.reshape(B, N, heads * C) + P[pre + "move_qkx_aft.bias"]
    qkx = qkx.reshape(B, N, heads, -1).permute(0, 2, 3, 1)
    attn = _einsum("BNC,BHCD -> BHND", xq, qkx) * (hd ** -0.5)
    if bias is not None:
        attn = attn + bias
    prob = _softmax_quant(attn, P, pre, abits)
    out = _matmul(prob, v).transpose(1, 2).reshape(B, N, C)
    _tap(pre + "@core_out", out=out)
    return qlinear(out, P, pre + "proj.", wbits, abits, symmetric=True)


# ------------------------------------------------------------------------------------------ 8-bit ends
def _lsq_generic(x: Tensor, alpha: Tensor, numel_per_scale: int, lo: int, hi: int) -> Tensor:
    g = 1.0 / ((hi * numel_per_scale) ** 0.5)
    return _lsq_core(x, alpha, g, lo, hi, 8, False)


def patch_embed_q(img: Tensor, P: Params, pre: str, state: dict) -> Tensor:
    """LSQ_QConv2d.forward (8/8 bit), qlinear.py:166-177, with LsqQuantizer4img (lsq.py:336-373),
    LearnableBias4img (qbias.py:20-23) and LsqQuantizer4Conv2d (lsq.py:419-437)."""
    w = P[pre + "weight"]
    lo, hi = -128, 127
    sw = _get_scale(P, pre + "lsqw_fn.s",
                    lambda: 2 * w.detach().abs().mean(dim=-1).mean(dim=-1).mean(dim=-1) / (hi ** 0.5))
    wq = _lsq_generic(w, sw.view(-1, 1, 1, 1), w.shape[1] * w.shape[2] * w.shape[3], lo, hi)
    x = img + P[pre + "move_b4.bias"].reshape(img.shape[-1], img.shape[-2]).expand_as(img)
    if float(x.detach().min()) < -1e-5:
        state["signed"] = 1  # sticky, lsq.py:338-339
    ilo, ihi = (0, 255) if not state.get("signed", 0) else (-128, 127)
    sx = _get_scale(P, pre + "input_quant_fn.s",
                    lambda: 2 * x.detach().abs().mean(dim=-1).mean(dim=-1).mean(dim=0) / (ihi ** 0.5))
    x = _lsq_generic(x, sx.view(1, -1, 1, 1), x.shape[0] * x.shape[2] * x.shape[3], ilo, ihi)
    x = x + P[pre + "move_aft.bias"].reshape(x.shape[-1], x.shape[-2]).expand_as(x)
    stride = w.shape[-1]
    return F.conv2d(x, wq, P[pre + "bias"], stride=stride)


def head_q(x: Tensor, P: Params, pre: str) -> Tensor:
    """LSQ_QLinear4head.forward (8/8 bit), qlinear.py:223-238, LsqQuantizerWeight (lsq.py:72-101) and
    LsqQuantizer4head_input (lsq.py:486-505)."""
    w = P[pre + "weight"]
    lo, hi = -128, 127
    sw = _get_scale(P, pre + "lsqw_fn.s", lambda: 2 * w.detach().abs().mean(dim=-1) / (hi ** 0.5))
    wq = _lsq_generic(w, sw.unsqueeze(-1), w.shape[-1], lo, hi)
    x = x + P[pre + "move_b4.bias"].expand_as(x)
    sx = _get_scale(P, pre + "input_quant_fn.s", lambda: (x.detach().abs().mean() * 2 / (hi ** 0.5)).reshape(1))
    x = _lsq_generic(x, sx, x.numel(), lo, hi)
    x = x + P[pre + "move_aft.bias"].expand_as(x)
    out = F.linear(x, wq)
    return out + P[pre + "bias"].view(1, -1).expand_as(out)


# ------------------------------------------------------------------------------------------ DeiT host
def deit_block(x: Tensor, P: Params, pre: str, heads: int, wbits: int, abits: int, qkr: bool) -> Tensor:
    """Block.forward, deit_vision_transformer.py:154-164 (LayerNorm eps 1e-6, deit.py:75)."""
    C = x.shape[-1]
    _tap(pre + "@in", x=x)
    h = F.layer_norm(x, (C,), P[pre + "norm1.weight"], P[pre + "norm1.bias"], 1e-6)
    attn = (qattention_qkr if qkr else qattention)(h, P, pre + "attn.", heads, wbits, abits)
    x = x + attn
    h = F.layer_norm(x, (C,), P[pre + "norm2.weight"], P[pre + "norm2.bias"], 1e-6)
    x = x + qmlp(h, P, pre + "mlp.", wbits, abits)
    _tap(pre + "@out", out=x)
    return x


def deit_forward(img: Tensor, P: Params, depth: int, heads: int, wbits: int, abits: int, qkr: bool,
                 state: Optional[dict] = None, training: bool = True):
    """DistilledVisionTransformer.forward, deit.py:32-67 with every qmodule of
    configs/ours_imagenet_recipe.attn_q.yml:47-74 quantized."""
    state = {} if state is None else state
    x = patch_embed_q(img, P, "patch_embed.proj.", state).flatten(2).transpose(1, 2)
    B = x.shape[0]
    x = torch.cat((P["cls_token"].expand(B, -1, -1), P["dist_token"].expand(B, -1, -1), x), dim=1)
    x = x + P["pos_embed"]
    for i in range(depth):
        x = deit_block(x, P, f"blocks.{i}.", heads, wbits, abits, qkr)
    C = x.shape[-1]
    x = F.layer_norm(x, (C,), P["norm.weight"], P["norm.bias"], 1e-6)
    cls = head_q(x[:, 0], P, "head.")
    dist = head_q(x[:, 1], P, "head_dist.")
    if training:
        return cls, dist
    return (cls + dist) / 2


# ------------------------------------------------------------------------------------------ Swin
def _rel_pos_index(ws):
    """src/swin.py:203-213."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), indexing="ij"))
    flat = torch.flatten(coords, 1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 0] *= 2 * ws[1] - 1
    return rel.sum(-1).view(-1)


def swin_window_attention(x: Tensor, P: Params, pre: str, heads: int, wbits: int, abits: int, qkr: bool,
                          window=(7, 7), shift=(0, 0)) -> Tensor:
    """QAttention_swin.forward / QAttention_swin_qkreparam.forward, swin_attention_and_mlp.py:143-251 / 344-461:
    pad, cyclic shift, window partition, quantized attention with relative-position bias and the 0/-100 shift mask,
    proj, reverse."""
    B, H, W, C = x.shape
    N = window[0] * window[1]
    bias = P[pre + "relative_position_bias_table"][_rel_pos_index(window)].view(N, N, -1).permute(2, 0, 1).contiguous().unsqueeze(0)
    pad_r = (window[1] - W % window[1]) % window[1]
    pad_b = (window[0] - H % window[0]) % window[0]
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    _, pH, pW, _ = x.shape
    shift = list(shift)
    if window[0] >= pH:
        shift[0] = 0
    if window[1] >= pW:
        shift[1] = 0
    if sum(shift) > 0:
        x = torch.roll(x, shifts=(-shift[0], -shift[1]), dims=(1, 2))
    nW = (pH // window[0]) * (pW // window[1])
    x = x.view(B, pH // window[0], window[0], pW // window[1], window[1], C).permute(0, 1, 3, 2, 4, 5).reshape(B * nW, N, C)
    full_bias = bias
    if sum(shift) > 0:
        m = x.new_zeros((pH, pW))
        hs = ((0, -window[0]), (-window[0], -shift[0]), (-shift[0], None))
        ws_ = ((0, -window[1]), (-window[1], -shift[1]), (-shift[1], None))
        cnt = 0
        for h in hs:
            for w in ws_:
                m[h[0]:h[1], w[0]:w[1]] = cnt
                cnt += 1
        m = m.view(pH // window[0], window[0], pW // window[1], window[1]).permute(0, 2, 1, 3).reshape(nW, N)
        m = m.unsqueeze(1) - m.unsqueeze(2)
        m = m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))
        full_bias = bias + m.repeat(B, 1, 1).unsqueeze(1)                      # [B*nW, 1|H, N, N]
    out = (qattention_qkr if qkr else qattention)(x, P, pre, heads, wbits, abits, bias=full_bias)
    out = out.view(B, pH // window[0], pW // window[1], window[0], window[1], C).permute(0, 1, 3, 2, 4, 5).reshape(B, pH, pW, C)
    if sum(shift) > 0:
        out = torch.roll(out, shifts=(shift[0], shift[1]), dims=(1, 2))
    return out[:, :H, :W, :].contiguous()


def swin_forward(img: Tensor, P: Params, depths, heads, wbits: int, abits: int, qkr: bool, state: Optional[dict] = None,
                 window=(7, 7)) -> Tensor:
    """SwinTransformer.forward (src/swin.py:430-448) with the qmodules of configs/swin_t_imagenet.attn_q.yml:44-73."""
    state = {} if state is None else state
    x = patch_embed_q(img, P, "features.0.0.", state).permute(0, 2, 3, 1)
    C = x.shape[-1]
    x = F.layer_norm(x, (C,), P["features.0.2.weight"], P["features.0.2.bias"], 1e-5)
    for i, depth in enumerate(depths):
        for j in range(depth):
            pre = f"features.{2 * i + 1}.{j}."
            C = x.shape[-1]
            _tap(pre + "@in", x=x)
            h = F.layer_norm(x, (C,), P[pre + "norm1.weight"], P[pre + "norm1.bias"], 1e-5)
            shift = (0, 0) if j % 2 == 0 else (window[0] // 2, window[1] // 2)
            x = x + swin_window_attention(h, P, pre + "attn.", heads[i], wbits, abits, qkr, window, shift)
            h = F.layer_norm(x, (C,), P[pre + "norm2.weight"], P[pre + "norm2.bias"], 1e-5)
            x = x + qmlp(h, P, pre + "mlp.", wbits, abits)
            _tap(pre + "@out", out=x)
        if i < len(depths) - 1:                                                 # PatchMerging, src/swin.py:40-59
            pre = f"features.{2 * i + 2}."
            Hh, Ww, _ = x.shape[-3:]
            x = F.pad(x, (0, 0, 0, Ww % 2, 0, Hh % 2))
            x = torch.cat([x[..., 0::2, 0::2, :], x[..., 1::2, 0::2, :], x[..., 0::2, 1::2, :], x[..., 1::2, 1::2, :]], -1)
            x = F.layer_norm(x, (x.shape[-1],), P[pre + "norm.weight"], P[pre + "norm.bias"], 1e-5)
            x = qlinear(x, P, pre + "reduction.", wbits, abits, symmetric=True)
    x = F.layer_norm(x, (x.shape[-1],), P["norm.weight"], P["norm.bias"], 1e-5)
    x = torch.flatten(F.adaptive_avg_pool2d(x.permute(0, 3, 1, 2), 1), 1)
    return head_q(x, P, "head.")


# ------------------------------------------------------------------------------------------ CGA
def cga_freeze_mask(w: Tensor, bits: int, boundary_range: float = 0.005) -> Tensor:
    """freeze_outside_boundary_weight_idx, cga.py:450-469: 1.0 where the weight is frozen (its pre-round
    StatsQ value is NOT within +-BR of a rounding boundary), 0.0 where it keeps training."""
    with torch.no_grad():
        b4, _ = statsq_pre_round(w, bits)
        r = torch.round(b4)
        lo, hi = int(r.min().item()), int(r.max().item())
        not_frozen = torch.zeros_like(w)
        for i in range(lo, hi):
            d = b4 - float(i)
            not_frozen = not_frozen + ((d <= (0.5 + boundary_range)) * (d >= (0.5 - boundary_range))).float()
        return 1.0 - not_frozen


def adamw_reference(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float,
                    beta2: float, eps: float, wd: float) -> None:
    """torch.optim.AdamW single-tensor math (torch/optim/adamw.py `_single_tensor_adamw`), in place."""
    with torch.no_grad():
        p.mul_(1 - lr * wd)
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        step_size = lr / bc1
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


def cga_masked_step(w: Tensor, grad: Tensor, m: Tensor, v: Tensor, step: int, lr: float, wd: float, bits: int,
                    boundary_range: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> Tensor:
    """cga.py:953-1013 for one weight: zero the gradient of frozen elements, stash them, AdamW, restore.
    Returns the freeze mask. Moments evolve as AdamW-with-zero-gradient for frozen elements."""
    with torch.no_grad():
        f = cga_freeze_mask(w, bits, boundary_range)
        g = grad * f * 0.0 + grad * (1 - f)
        stash = (w * f).clone()
        adamw_reference(w, g, m, v, step, lr, beta1, beta2, eps, wd)
        w.copy_(w * (1 - f) + stash)
        return f


# ----------------------------------------------------------------------------------------------- KD losses (f2)
def kl_loss_soft(output, target, T: float = 1.0, reduction: str = "mean"):
    """KLLossSoft.forward (src/quantization/utils.py:44-58): soft-target cross entropy -sum softmax(t/T) * log_softmax(o/T);
    tuples (the distilled student / the training-mode teacher) contribute their FIRST element."""
    output = output[0] if isinstance(output, tuple) else output
    target = target[0] if isinstance(target, tuple) else target
    output, target = output / T, target / T
    loss = -torch.sum(torch.softmax(target, dim=1) * torch.log_softmax(output, dim=1), dim=1)
    return loss.mean() if reduction == "mean" else (loss.sum() if reduction == "sum" else loss)


def kd_loss_soft_and_hard(output, hard_target, soft_target):
    """KDLossSoftandHard.forward (src/quantization/utils.py:60-77): hard CE on the class head + soft CE of the distillation head
    against the teacher (both on the same logits for a single-output student)."""
    if isinstance(output, tuple):
        return kl_loss_soft(output[1], soft_target) + torch.nn.functional.cross_entropy(output[0], hard_target)
    return kl_loss_soft(output, soft_target) + torch.nn.functional.cross_entropy(output, hard_target)
