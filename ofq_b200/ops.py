"""Tensor-level wrappers over the C-ABI (include/ofq_b200.h).

Each function checks device / dtype / contiguity, allocates outputs with torch (plumbing) and launches the
sm_100a kernel on torch's current stream.  No function here computes anything itself.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import GemmLsq, GemmOut, Operand, Vec, check

GEMM_I8, GEMM_BF16, GEMM_F16 = 0, 1, 2
FMT_BF16, FMT_F16 = 0, 1
PER_ROW, PER_COL = 0, 1
_T16 = {FMT_BF16: torch.bfloat16, FMT_F16: torch.float16}

# Launch accounting (bench.py): LAUNCHES counts kernels launched through this module; when PROFILE is a list every
# call is bracketed by CUDA events on the launching stream and recorded as
# (family, start_event, end_event, algorithmic_bytes, algorithmic_flops).
LAUNCHES = 0
PROFILE = None
TAG = ""          # set by callers that want the next launches attributed to a site (bench.py per-site table)


def _call(family: str, nkernels: int, alg_bytes: float, alg_flops: float, fn, *args, tag: str = ""):
    global LAUNCHES
    LAUNCHES += nkernels
    if PROFILE is None:
        check(fn(*args))
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    check(rc)
    PROFILE.append((family, e0, e1, alg_bytes, alg_flops, tag))


# ------------------------------------------------------------------------------------------------ side stream
# Weight-gradient work (the split-K dW GEMMs, the W_qk product backward) does not feed the activation-gradient chain of a
# layer: inside one autograd backward it can be queued on a second stream and joined before the backward returns (so
# autograd, DDP and the optimizer only ever see completed gradients on the current stream; works the same under CUDA-graph
# capture: fork / join edges). OFF by default (OFQ_SIDE_STREAM=1 turns it on): measured on B200, the persistent GEMM CTAs
# (168 registers x 320 threads, ~200 KB of shared memory) leave no room for a streaming CTA on the same SM, so the two
# streams only take SMs from each other: 22.9 ms per step with the side stream against 22.3 ms without.
SIDE_ENABLED = os.environ.get("OFQ_SIDE_STREAM", "0") == "1"
_SIDE = {}
_SIDE_KEEP = []


class side_stream:
    """`with ops.side_stream(t0, t1, ...):` launches issued inside run on the side stream after everything already queued
    on the current stream; the tensors named are kept alive until `side_join()` (the caching allocator must not hand
    their storage to later work of the current stream while the side stream still reads it)."""

    def __init__(self, *keep):
        self.keep = keep
        self.ctx = None

    def __enter__(self):
        if not SIDE_ENABLED:
            return self
        cur = torch.cuda.current_stream()
        side = _SIDE.get(cur.device)
        if side is None:
            # higher priority: when both become runnable the one-CTA-per-SM GEMM is placed first and the streaming
            # kernel of the main chain fills the rest of each SM (the other order would serialise them)
            side = _SIDE[cur.device] = torch.cuda.Stream(device=cur.device, priority=int(os.environ.get("OFQ_SIDE_PRIO", "-1")))
        side.wait_stream(cur)
        _SIDE_KEEP.append((cur, side, self.keep))
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def side_join() -> None:
    """The current stream waits for everything forked with `side_stream` since the last join."""
    if not _SIDE_KEEP:
        return
    cur, side, _ = _SIDE_KEEP[-1]
    cur.wait_stream(side)
    _SIDE_KEEP.clear()


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.OfqError("ofq_b200 ops need CUDA tensors (there is no CPU fallback)")


def _st():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------ zero arena
# The backward of every layer needs a few tiny pre-zeroed scratch vectors (running maxima for the fp16 range scales, scale
# quadruples, atomic accumulators): as `torch.zeros` each is a fill launch of its own - ~70 per DeiT-S step. They are carved
# from ONE persistent buffer per device instead, zeroed by a single memset when the step begins (StepPrologue.begin).
# Only for scratch that lives and dies inside one autograd-node call: nothing carved here may be returned as a gradient or
# saved for the backward (the next step's memset would clear it).
_ARENA = {}
_ARENA_FLOATS = 1 << 14


class _ZeroArena:
    __slots__ = ("buf", "cursor", "live", "carved", "missed")

    def __init__(self, device):
        self.buf = torch.zeros(_ARENA_FLOATS, dtype=torch.float32, device=device)
        self.cursor, self.live, self.carved, self.missed = 0, False, 0, 0


def arena_begin(device) -> None:
    """Start of a step on `device`: one memset clears every scratch vector the step will carve."""
    a = _ARENA.get(device)
    if a is None:
        a = _ARENA[device] = _ZeroArena(device)
    a.buf.zero_()
    a.cursor, a.live = 0, True


def scratch_zeros(n: int, device, dtype=torch.float32) -> torch.Tensor:
    """`n` zeroed 4-byte elements of call-local scratch (see above); falls back to torch.zeros outside a step or when the
    arena is exhausted."""
    a = _ARENA.get(device)
    need = (n + 3) & ~3                                  # keep every carve 16-byte aligned
    if a is None or not a.live or a.cursor + need > _ARENA_FLOATS:
        if a is not None:
            a.missed += 1
        return torch.zeros(n, dtype=dtype, device=device)
    t = a.buf[a.cursor:a.cursor + n]
    a.cursor += need
    a.carved += 1
    return t if dtype == torch.float32 else t.view(dtype)


def _ptr(t):
    return None if t is None else t.data_ptr()


def vec(t: Optional[torch.Tensor], period: int = 0, bs1: int = 0, bs2: int = 0):
    if t is None:
        return None
    assert t.dtype == torch.float32
    return C.byref(Vec(t.data_ptr(), period, bs1, bs2))


def gemm(kind: int, a: torch.Tensor, a_strides, b: torch.Tensor, b_strides, out: torch.Tensor, out_strides,
         M: int, N: int, K: int, *, k2: int = 1, nb1: int = 1, nb2: int = 1, splits: int = 1, accumulate: bool = False,
         rs=None, cs=None, rt=None, ct=None, a_k2mod: int = 0, b_k2mod: int = 0, a_dual_delta: int = 0,
         a_mn: bool = False, b_mn: bool = False, amax: Optional[torch.Tensor] = None) -> None:
    """Raw ofq_gemm call. a_strides/b_strides = (row, k2, batch1, batch2) in elements; out_strides = (ld, b1, b2).
    a_mn / b_mn: the operand is stored [k][row] (MN-major) and its first stride is the distance between consecutive k.
    rs/cs/rt/ct are ctypes Vec references from `vec()` or None. amax: one fp32 element pre-set to 0 that receives max |out|."""
    _cuda(a, b, out)
    A = Operand(a.data_ptr(), a_strides[0], a_strides[1], a_k2mod, a_dual_delta, a_strides[2], a_strides[3], int(a_mn))
    B = Operand(b.data_ptr(), b_strides[0], b_strides[1], b_k2mod, 0, b_strides[2], b_strides[3], int(b_mn))
    O = GemmOut(out.data_ptr(), *out_strides, 1 if accumulate else 0)
    eb = 1 if kind == GEMM_I8 else 2
    tag = f"M{M} N{N} K{K} k2={k2}{'+dual' if a_dual_delta else ''} nb={nb1}x{nb2} sp={splits}{' aT' if a_mn else ''}{' bT' if b_mn else ''}"

    def _n(strides, mod):
        k2n = ((min(mod, k2) if mod else k2) + (a_dual_delta if strides is a_strides and a_dual_delta else 0)) if strides[1] else 1
        return k2n * (nb1 if strides[2] else 1) * (nb2 if strides[3] else 1)
    nout = (nb1 if out_strides[1] else 1) * (nb2 if out_strides[2] else 1)
    alg_bytes = eb * K * (M * _n(a_strides, a_k2mod) + N * _n(b_strides, b_k2mod)) + 4 * M * N * nout * (2 if accumulate else 1)
    alg_flops = 2.0 * M * N * K * k2 * nb1 * nb2 * (2 if a_dual_delta else 1)
    _call(("gemm_i8", "gemm_bf16", "gemm_f16")[kind], 1, alg_bytes, alg_flops, _lib.load().ofq_gemm_ex, kind,
          C.byref(A), C.byref(B), C.byref(O), M, N, K, k2, nb1, nb2, splits, rs, cs, rt, ct, _ptr(amax), _st(), tag=tag)


def gemm_lsq(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, b4: Optional[torch.Tensor], s2: torch.Tensor, period: int,
             nseg: int, qlo: int, qhi: int, *, rs=None, cs=None, rt=None, ct=None, fmt16: Optional[int] = None,
             want_res: bool = False, dot_u: Optional[torch.Tensor] = None):
    """int8 GEMM a [M, K] x b [N, K]^T whose epilogue is the LSQ quantizer of the product (ofq_gemm_lsq): the fp32 output never
    exists. s2 = [s_eff, 1 / s_eff] with index (m % period) * nseg + n // (N // nseg). Returns (codes int8 [M, N],
    codes16 | None, res16 | None, rowdot [M, nseg] | None); res16 is the fp16 residual plane lsq_bwd(act=ACT_RES16) reads."""
    _cuda(a, b, s2)
    assert a.dtype == torch.int8 and b.dtype == torch.int8 and a.stride(1) == 1 and b.stride(1) == 1 and N % nseg == 0
    dev = a.device
    codes = torch.empty((M, N), dtype=torch.int8, device=dev)
    c16 = torch.empty((M, N), dtype=_T16[fmt16], device=dev) if fmt16 is not None else None
    res = torch.empty((M, N), dtype=torch.float16, device=dev) if want_res else None
    rowdot = torch.empty((M, nseg), dtype=torch.float32, device=dev) if dot_u is not None else None
    ws = torch.empty((M, (N + 31) // 32), dtype=torch.float32, device=dev) if dot_u is not None else None
    A = Operand(a.data_ptr(), a.stride(0), 0, 0, 0, 0, 0, 0)
    B = Operand(b.data_ptr(), b.stride(0), 0, 0, 0, 0, 0, 0)
    q = GemmLsq(codes.data_ptr(), N, _ptr(c16), N, fmt16 if fmt16 is not None else FMT_F16, _ptr(res), N, _ptr(b4), s2[0].data_ptr(),
                s2[1].data_ptr(), period, nseg, N // nseg, float(qlo), float(qhi), _ptr(dot_u), _ptr(rowdot), _ptr(ws))
    nbytes = float(M * K + N * K) + M * N * (1.0 + (2 if c16 is not None else 0) + (2 if res is not None else 0))
    _call("gemm_lsq", 2 if dot_u is not None else 1, nbytes, 2.0 * M * N * K, _lib.load().ofq_gemm_lsq, C.byref(A), C.byref(B), M, N, K,
          rs, cs, rt, ct, C.byref(q), _st(), tag=f"M{M} N{N} K{K} lsq")
    return codes, c16, res, rowdot


def gemm_dx_lsq(kind: int, a16: torch.Tensor, a_strides, b16: torch.Tensor, b_strides, M: int, N: int, K: int, *, rs, cs, x2d: torch.Tensor,
                b4: torch.Tensor, period: int, qlo: int, qhi: int, g: float, w_codes: torch.Tensor, dy_colsum: torch.Tensor,
                colscale: Optional[torch.Tensor] = None, a_mn: bool = False, b_mn: bool = False, amax: Optional[torch.Tensor] = None):
    """dX GEMM of a quantized linear layer with the LSQ backward of its input quantizer as the epilogue (ofq_gemm_dx_lsq):
    returns (dx [M, N] fp32, d_s [min(period, M)], d_b4 [N], d_aft [N]); dX_hat itself is never written. rs = vec(1 / s_eff, period)
    is both the GEMM's row un-scale and the quantizer's reciprocal step; cs the period-1 range un-scale. d_aft comes from the
    identity sum_m dX_hat = (colsum(dY) * colscale) . w_codes (w_codes int8 [K, N]: the layer's weight codes)."""
    _cuda(a16, b16, x2d, b4)
    assert x2d.dtype == torch.float32 and x2d.stride(1) == 1 and tuple(x2d.shape) == (M, N)
    dev = x2d.device
    lib = _lib.load()
    dx = torch.empty((M, N), dtype=torch.float32, device=dev)
    d_s = torch.empty(min(period, M), dtype=torch.float32, device=dev)
    d_b4 = torch.empty(N, dtype=torch.float32, device=dev)
    d_aft = torch.empty(N, dtype=torch.float32, device=dev)
    ws = torch.empty(lib.ofq_gemm_dx_lsq_workspace(M, N), dtype=torch.float32, device=dev)
    A = Operand(a16.data_ptr(), a_strides[0], 0, 0, 0, 0, 0, int(a_mn))
    B = Operand(b16.data_ptr(), b_strides[0], 0, 0, 0, 0, 0, int(b_mn))
    assert w_codes.dtype == torch.int8 and tuple(w_codes.shape) == (K, N) and w_codes.stride(1) == 1 and dy_colsum.numel() == K
    _call("gemm_dx_lsq", 3, 2.0 * (M * K + N * K) + 8.0 * M * N, 2.0 * M * N * K, lib.ofq_gemm_dx_lsq, kind, C.byref(A), C.byref(B), M, N, K,
          rs, cs, x2d.data_ptr(), x2d.stride(0), b4.data_ptr(), qlo, qhi, float(g), dx.data_ptr(), N, d_s.data_ptr(), d_b4.data_ptr(),
          d_aft.data_ptr(), w_codes.data_ptr(), w_codes.stride(0), dy_colsum.data_ptr(), _ptr(colscale), _ptr(amax), ws.data_ptr(), _st(),
          tag=f"M{M} N{N} K{K} dx+lsq")
    return dx, d_s, d_b4, d_aft


# ------------------------------------------------------------------------------------------------ quantizers
def statsq_codes(w: torch.Tensor, bits: int, aft: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                 want_minmax: bool = False, want_inv: bool = False, want_sf: bool = False, fmt16: Optional[int] = None):
    """StatsQ codes of a 2-D fp32 weight. Returns (codes int8 [R,C], colscale [R], sf [R] | None, colterm [R] | None,
    kminmax int32[2] | None [, 1/colscale [R] if want_inv]). sf is only guaranteed with want_sf (the step prologue does
    not produce it). fmt16: a further last element, the exact 16-bit copy [R,C] of the codes."""
    _cuda(w)
    assert w.dim() == 2 and w.dtype == torch.float32 and w.stride(1) == 1
    R, Cc = w.shape
    from . import prologue
    pro = prologue.ACTIVE if prologue.ENABLED else None
    c16 = None
    if pro is not None and not want_minmax and not want_sf:
        job, ready = pro.get_statsq(w, bits, aft, bias, fmt16)   # persistent outputs; `ready`: produced by the step prologue
        codes, cs2, colterm, c16 = job.out["codes"], job.out["cs2"], job.out["colterm"], job.out["codes16"]
        if ready:
            res = (codes, cs2[0], None, colterm, None) + ((cs2[1],) if want_inv else ())
            return res + ((c16,) if fmt16 is not None else ())
    else:
        codes = torch.empty((R, Cc), dtype=torch.int8, device=w.device)
        cs2 = torch.empty((2, R), dtype=torch.float32, device=w.device)
        colterm = torch.empty(R, dtype=torch.float32, device=w.device) if (aft is not None or bias is not None) else None
        c16 = torch.empty((R, Cc), dtype=_T16[fmt16], device=w.device) if fmt16 is not None else None
    colscale = cs2[0]
    sf = torch.empty(R, dtype=torch.float32, device=w.device)
    mm = None
    if want_minmax:
        mm = torch.tensor([2 ** 31 - 1, -2 ** 31], dtype=torch.int32, device=w.device)
    if colterm is not None and aft is None:
        aft = torch.zeros(Cc, dtype=torch.float32, device=w.device)
    _call("statsq", 1, 5.0 * R * Cc, 0, _lib.load().ofq_statsq_codes_ex, w.data_ptr(), R, Cc, w.stride(0), bits,
          codes.data_ptr(), Cc, colscale.data_ptr(), sf.data_ptr(), _ptr(aft), _ptr(bias), _ptr(colterm), _ptr(mm),
          cs2[1].data_ptr(), _ptr(c16), fmt16 if fmt16 is not None else FMT_F16, _st())
    res = (codes, colscale, sf, colterm, mm) + ((cs2[1],) if want_inv else ())
    return res + ((c16,) if fmt16 is not None else ())


def lsq_effective_scale(alpha: torch.Tensor, g: float, recip: bool = False):
    """Effective LSQ step sizes; recip=True returns [2, n]: row 0 = scales, row 1 = their reciprocals."""
    _cuda(alpha)
    a = alpha.detach().contiguous()
    from . import prologue
    pro = prologue.ACTIVE if prologue.ENABLED else None
    if pro is not None and a.dtype == torch.float32:
        job, ready = pro.get_scale(a, g, recip)
        if not ready:
            _call("lsq_scale", 1, (12.0 if recip else 8.0) * a.numel(), 0, _lib.load().ofq_lsq_effective_scale, a.data_ptr(),
                  a.numel(), float(g), (job.out[0] if recip else job.out).data_ptr(), job.out[1].data_ptr() if recip else None, _st())
        return job.out
    if recip:
        out = torch.empty((2,) + tuple(a.shape), dtype=a.dtype, device=a.device)
        _call("lsq_scale", 1, 12.0 * a.numel(), 0, _lib.load().ofq_lsq_effective_scale, a.data_ptr(), a.numel(), float(g),
              out[0].data_ptr(), out[1].data_ptr(), _st())
        return out
    out = torch.empty_like(a)
    _call("lsq_scale", 1, 8.0 * a.numel(), 0, _lib.load().ofq_lsq_effective_scale, a.data_ptr(), a.numel(), float(g),
          out.data_ptr(), None, _st())
    return out


ACT_NONE, ACT_GELU, ACT_RES16 = 0, 1, 2      # ACT_RES16 (lsq_bwd with out16 only): x2d is gemm_lsq's fp16 residual plane


def lsq_quant(x2d: torch.Tensor, b4: torch.Tensor, s_eff: torch.Tensor, mode: int, period: int, nseg: int,
              qlo: int, qhi: int, out: Optional[torch.Tensor] = None, act: int = ACT_NONE, fmt16: Optional[int] = None,
              dot_u: Optional[torch.Tensor] = None):
    """x2d: [rows, cols] fp32 view (last dim contiguous). Returns int8 codes [rows, cols] of Q(act(x) + b4); with fmt16
    (FMT_BF16 / FMT_F16) returns (codes, exact 16-bit copy [rows, cols]) written in the same pass; with dot_u [cols] a
    further element: segment-wise dot products of the codes with dot_u, [rows, nseg] (codes_rowdot in the same pass)."""
    _cuda(x2d, b4, s_eff)
    assert x2d.dim() == 2 and x2d.stride(1) == 1 and x2d.dtype == torch.float32
    rows, cols = x2d.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.int8, device=x2d.device)
    if act == ACT_NONE and fmt16 is None and dot_u is None:
        _call("lsq_quant", 1, 5.0 * rows * cols, 0, _lib.load().ofq_lsq_quant, x2d.data_ptr(), rows, cols, x2d.stride(0),
              b4.data_ptr(), s_eff.data_ptr(), mode, period, nseg, qlo, qhi, out.data_ptr(), out.stride(0), _st())
        return out
    out16 = torch.empty((rows, cols), dtype=_T16[fmt16], device=x2d.device) if fmt16 is not None else None
    fused_dot = dot_u is not None and cols % 128 == 0 and (nseg == 1 or (cols // nseg) % 128 == 0) and x2d.stride(0) % 4 == 0
    part = torch.empty(((cols // nseg) // 128, rows, nseg), dtype=torch.float32, device=x2d.device) if fused_dot else None
    _call("lsq_quant", 1, (5.0 + (2 if fmt16 is not None else 0)) * rows * cols, 0, _lib.load().ofq_lsq_quant_ex, x2d.data_ptr(),
          rows, cols, x2d.stride(0), b4.data_ptr(), s_eff.data_ptr(), mode, period, nseg, qlo, qhi, act, out.data_ptr(),
          out.stride(0), _ptr(out16), cols, fmt16 if fmt16 is not None else FMT_F16, _ptr(dot_u) if fused_dot else None,
          _ptr(part), _st())
    res = (out,) if fmt16 is None else (out, out16)
    if dot_u is not None:
        # the partial planes are summed in a fixed order (deterministic logits); unfused shapes take the stand-alone kernel
        res += ((part.sum(0) if part.shape[0] > 1 else part[0]) if fused_dot else codes_rowdot(out, nseg, dot_u),)
    return res[0] if len(res) == 1 else res


def lsq_bwd(dy2d: torch.Tensor, x2d: torch.Tensor, b4: torch.Tensor, s_eff: torch.Tensor, mode: int, period: int,
            nseg: int, qlo: int, qhi: int, g: float, want_ds: bool = True, want_aft: bool = True, next_scale=None,
            zero_sum: bool = False, act: int = ACT_NONE, out16=None, want_dx: bool = True, want_colsum: bool = False):
    """Returns (dx [rows, cols], d_s, d_b4 [cols], d_aft [cols] | None).  next_scale = (v1, v2, mult, product): additionally
    returns the fp16 range scales (absmax_scale layout) of dx*v1[c] / dx*v2[r] for the GEMM operand made from dx,
    derived from max|dx| at no extra pass over dx.  act: the quantizer saw act(x2d) (lsq_quant(act=...)); dx is then the
    gradient w.r.t. the pre-activation x2d.  out16 = (fmt, cs [cols], rs, rs_period, scale4): the same pass also writes the
    16-bit GEMM operand rn16(dx * cs[c] * rs[r % rs_period] * scale4[0]) (appended to the result); with want_dx=False the
    fp32 dx is not written at all (returned as None); want_colsum (per-row scale mode, with out16) appends colsum(dx)."""
    _cuda(dy2d, x2d)
    assert dy2d.dim() == 2 and dy2d.stride(1) == 1 and x2d.stride(1) == 1
    rows, cols = dy2d.shape
    lib = _lib.load()
    ws = torch.empty(lib.ofq_lsq_bwd_workspace(rows, cols, nseg), dtype=torch.float32, device=dy2d.device)
    dx = torch.empty((rows, cols), dtype=torch.float32, device=dy2d.device) if (want_dx or out16 is None) else None
    o16 = None
    if out16 is None:
        _call("lsq_bwd", 1, 12.0 * rows * cols, 0, lib.ofq_lsq_bwd_act, dy2d.data_ptr(), dy2d.stride(0), x2d.data_ptr(),
              x2d.stride(0), rows, cols, b4.data_ptr(), s_eff.data_ptr(), mode, period, nseg, qlo, qhi, act, dx.data_ptr(),
              dx.stride(0), ws.data_ptr(), _st())
    else:
        fmt16, cs16, rs16, rs16_period, scale4 = out16
        o16 = torch.empty((rows, cols), dtype=_T16[fmt16], device=dy2d.device)
        _call("lsq_bwd", 1, (10.0 + (4 if dx is not None else 0)) * rows * cols, 0, lib.ofq_lsq_bwd_ex, dy2d.data_ptr(),
              dy2d.stride(0), x2d.data_ptr(), x2d.stride(0), rows, cols, b4.data_ptr(), s_eff.data_ptr(), mode, period, nseg, qlo,
              qhi, act, _ptr(dx), cols, o16.data_ptr(), cols, fmt16, _ptr(cs16), _ptr(rs16), rs16_period, _ptr(scale4),
              ws.data_ptr(), _st())
    ns = cols if mode == PER_COL else min(period, rows) * nseg
    d_s = torch.empty(ns, dtype=torch.float32, device=dy2d.device) if want_ds else None
    d_b4 = torch.empty(cols, dtype=torch.float32, device=dy2d.device)
    d_aft = torch.empty(cols, dtype=torch.float32, device=dy2d.device) if want_aft else None
    if next_scale is None:
        csum = torch.empty(cols, dtype=torch.float32, device=dy2d.device) if (want_colsum and out16 is not None) else None
        _call("lsq_bwd_finalize", 1, 4.0 * ws.numel(), 0, lib.ofq_lsq_bwd_finalize_colsum, ws.data_ptr(), rows, cols, mode, period,
              nseg, float(g), _ptr(d_s), d_b4.data_ptr(), _ptr(d_aft), int(zero_sum), _ptr(csum), _st())
        if out16 is not None:
            return (dx, d_s, d_b4, d_aft, o16, csum) if csum is not None else (dx, d_s, d_b4, d_aft, o16)
        return dx, d_s, d_b4, d_aft
    v1, v2, mult, product = next_scale
    sc = torch.empty(4, dtype=torch.float32, device=dy2d.device)
    _call("lsq_bwd_finalize", 1, 4.0 * ws.numel(), 0, lib.ofq_lsq_bwd_finalize_scale, ws.data_ptr(), rows, cols, mode, period,
          nseg, float(g), _ptr(d_s), d_b4.data_ptr(), _ptr(d_aft), int(zero_sum), _ptr(v1), 0 if v1 is None else v1.numel(),
          _ptr(v2), 0 if v2 is None else v2.numel(), float(mult), int(product), sc.data_ptr(), _st())
    return dx, d_s, d_b4, d_aft, sc


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def scale_from_max(amax: torch.Tensor, *, v1=None, v2=None, mult: float = 1.0, product: bool = False) -> torch.Tensor:
    """fp16 range scales (absmax_scale layout) from maxima tracked elsewhere (e.g. gemm(amax=...))."""
    out = torch.empty(4, dtype=torch.float32, device=amax.device)
    _call("absmax_scale", 1, 0.0, 0, _lib.load().ofq_scale_from_max, amax.data_ptr(), amax.numel(), _ptr(v1),
          0 if v1 is None else v1.numel(), _ptr(v2), 0 if v2 is None else v2.numel(), float(mult), int(product), out.data_ptr(),
          _st())
    return out


_ABSMAX_WS = {}


def absmax_scale(x: torch.Tensor, nb: int, R: int, Cc: int, ldx: int, bstride: int, *, cs=None, rs=None, rs_period=0,
                 v1=None, v2=None, mult: float = 1.0, product: bool = False) -> torch.Tensor:
    """Power-of-two fp16 range scales of x*cs (out[0], inverse out[1]) and x*rs (out[2], inverse out[3]); the bounds are
    further multiplied by max|v1| / max|v2| and `mult` (see ofq_absmax_scale)."""
    _cuda(x)
    lib = _lib.load()
    ws = _ABSMAX_WS.get(x.device)
    if ws is None:
        ws = _ABSMAX_WS[x.device] = torch.zeros(lib.ofq_absmax_scale_workspace(), dtype=torch.int32, device=x.device)
    out = torch.empty(4, dtype=torch.float32, device=x.device)
    _call("absmax_scale", 1, 4.0 * nb * R * Cc, 0, lib.ofq_absmax_scale, x.data_ptr(), nb, R, Cc, ldx, bstride, _ptr(cs),
          _ptr(rs), rs_period, _ptr(v1), 0 if v1 is None else v1.numel(), _ptr(v2), 0 if v2 is None else v2.numel(),
          float(mult), int(product), out.data_ptr(), ws.data_ptr(), _st())
    return out


def grad_prep(x: torch.Tensor, nb: int, R: int, Cc: int, ldx: int, bstride: int, *, cs=None, rs=None, rs_period=0,
              want_rm: bool = False, want_t: bool = False, want_colsum: bool = False, u=None, group: int = 64,
              planes: int = 1, fmt: int = FMT_BF16, scale4=None, rm_rowscale: bool = False):
    """One pass over a fp32 gradient: returns dict(rm=16-bit [planes,nb,R,C], t=16-bit [planes,nb,C,r_pad], r_pad,
    colsum [C], rowdot [nb,C/group,R])."""
    _cuda(x)
    dev = x.device
    r_pad = round_up(R, 8)
    out = {"r_pad": r_pad}
    rm = torch.empty((planes, nb, R, Cc), dtype=_T16[fmt], device=dev) if want_rm else None
    t = torch.empty((planes, nb, Cc, r_pad), dtype=_T16[fmt], device=dev) if want_t else None
    colsum = torch.zeros(Cc, dtype=torch.float32, device=dev) if want_colsum else None
    rowdot = torch.empty((nb, Cc // group, R), dtype=torch.float32, device=dev) if u is not None else None
    nbytes = nb * R * Cc * (4.0 + 2 * planes * (int(want_rm) + int(want_t)))
    _call("grad_prep", 1, nbytes, 0, _lib.load().ofq_grad_prep, x.data_ptr(), nb, R, Cc, ldx, bstride, _ptr(cs), _ptr(rs),
          rs_period, planes, _ptr(rm), Cc, _ptr(t), r_pad, _ptr(colsum), _ptr(u), group, _ptr(rowdot), fmt, _ptr(scale4), int(rm_rowscale), _st())
    out.update(rm=rm, t=t, colsum=colsum, rowdot=rowdot)
    return out


def codes_to_bf16(codes: torch.Tensor, nb: int, R: int, Cc: int, ld: int, bstride: int, transpose: bool,
                  fmt: int = FMT_BF16):
    """int8 codes [nb][R][C] -> bf16 / fp16 (exact) [nb,R,C] or transposed [nb,C,r_pad]."""
    _cuda(codes)
    if transpose:
        r_pad = round_up(R, 8)
        out = torch.empty((nb, Cc, r_pad), dtype=_T16[fmt], device=codes.device)
        _call("codes_to_bf16", 1, 3.0 * nb * R * Cc, 0, _lib.load().ofq_codes_to_16, codes.data_ptr(), nb, R, Cc, ld,
              bstride, out.data_ptr(), r_pad, Cc * r_pad, 1, fmt, _st())
    else:
        out = torch.empty((nb, R, Cc), dtype=_T16[fmt], device=codes.device)
        _call("codes_to_bf16", 1, 3.0 * nb * R * Cc, 0, _lib.load().ofq_codes_to_16, codes.data_ptr(), nb, R, Cc, ld,
              bstride, out.data_ptr(), Cc, R * Cc, 0, fmt, _st())
    return out


def codes_transpose(codes: torch.Tensor, nb: int, R: int, Cc: int, ld: int, bstride: int):
    """int8 codes [nb][R][C] -> int8 [nb, C, r_pad16] (zero padded)."""
    _cuda(codes)
    r_pad = round_up(R, 16)
    out = torch.empty((nb, Cc, r_pad), dtype=torch.int8, device=codes.device)
    _call("codes_transpose", 1, 2.0 * nb * R * Cc, 0, _lib.load().ofq_codes_transpose, codes.data_ptr(), nb, R, Cc, ld,
          bstride, out.data_ptr(), r_pad, Cc * r_pad, _st())
    return out


def codes_rowdot(codes2d: torch.Tensor, nseg: int, u: torch.Tensor) -> torch.Tensor:
    """[rows, cols] int8 codes x fp32 vector u[cols] -> [rows, nseg] segment-wise dot products."""
    _cuda(codes2d, u)
    rows, cols = codes2d.shape
    out = torch.empty((rows, nseg), dtype=torch.float32, device=codes2d.device)
    _call("codes_rowdot", 1, 1.0 * rows * cols, 0, _lib.load().ofq_codes_rowdot, codes2d.data_ptr(), rows, cols,
          codes2d.stride(0), nseg, u.data_ptr(), out.data_ptr(), _st())
    return out


# ------------------------------------------------------------------------------------------------ softmax
def softmax_quant(S: torch.Tensor, N: int, H: int, s_eff: torch.Tensor, qhi: int, *, bias=None, mask=None, nW: int = 0,
                  save_p: bool = True, fmt16: Optional[int] = None):
    """S: [nz, N, ld] fp32 scaled logits. Returns (P fp32 [nz,N,ld] | None, codes int8 [nz,N,ldq], rowsum [nz,N]) and, with
    fmt16 (no bias / mask), a fourth element: the exact 16-bit copy of the codes [nz,N,ldq]."""
    _cuda(S, s_eff)
    nz, _, ld = S.shape
    ldq = round_up(N, 16)
    P = torch.empty_like(S) if save_p else None
    codes = torch.empty((nz, N, ldq), dtype=torch.int8, device=S.device)
    rowsum = torch.empty((nz, N), dtype=torch.float32, device=S.device)
    c16 = torch.empty((nz, N, ldq), dtype=_T16[fmt16], device=S.device) if fmt16 is not None else None
    _call("softmax_quant", 1, nz * N * N * (5.0 + (4 if save_p else 0) + (2 if fmt16 is not None else 0)), 0,
          _lib.load().ofq_softmax_quant_ex, S.data_ptr(), nz, N, ld, H, _ptr(bias), _ptr(mask), nW, s_eff.data_ptr(), qhi,
          _ptr(P), codes.data_ptr(), ldq, rowsum.data_ptr(), _ptr(c16), fmt16 if fmt16 is not None else FMT_F16, _st())
    if fmt16 is not None:
        return P, codes, rowsum, c16
    return P, codes, rowsum


def softmax_quant_bwd(dPq: torch.Tensor, P: torch.Tensor, N: int, H: int, s_eff: torch.Tensor, qhi: int, alpha: float,
                      g_s: float, ca: torch.Tensor, ca_per_head: bool, rb: torch.Tensor, want_ds32: bool = False,
                      planes: int = 1, fmt: int = FMT_BF16, scale4=None, single: bool = False):
    """single: only out_a = dS * ca[d] * rb[n] is produced (out_bt is None).
    Returns (out_a 16-bit [B,planes,H,N,ldo], out_bt [B,planes,H,N,ldo] | None, ldo, colsum [nz,N], d_s [N], dS32 | None)."""
    _cuda(dPq, P)
    nz, _, ld = P.shape
    ldo = round_up(N, 8)
    dev = P.device
    out_a = torch.empty((nz // H, planes, H, N, ldo), dtype=_T16[fmt], device=dev)
    out_bt = None if single else torch.empty((nz // H, planes, H, N, ldo), dtype=_T16[fmt], device=dev)
    # the vectorised path (single output, one plane, aligned rows) writes colsum and d_s; the generic one accumulates into them
    vec_path = single and planes == 1 and ld % 4 == 0
    colsum = torch.zeros((nz, N), dtype=torch.float32, device=dev)
    d_s = torch.zeros(N, dtype=torch.float32, device=dev)
    ds_part = torch.empty((nz, N), dtype=torch.float32, device=dev) if vec_path else None
    ds32 = torch.empty_like(P) if want_ds32 else None
    _call("softmax_quant_bwd", 2 if vec_path else 1, nz * N * N * (8.0 + (2 if single else 4) * planes + (4 if want_ds32 else 0)), 0,
          _lib.load().ofq_softmax_quant_bwd_ex, dPq.data_ptr(), P.data_ptr(), nz, N, ld, H, s_eff.data_ptr(), qhi,
          float(alpha), float(g_s), _ptr(ca), 1 if ca_per_head else 0, _ptr(rb), planes, out_a.data_ptr(),
          _ptr(out_bt), ldo, colsum.data_ptr(), d_s.data_ptr(), _ptr(ds32), fmt, _ptr(scale4), int(single), _ptr(ds_part), _st())
    return out_a, out_bt, ldo, colsum, d_s, ds32


def qkr_attn_fwd(qx: torch.Tensor, qk: torch.Tensor, qvT: torch.Tensor, B: int, N: int, H: int, Cc: int, se_x, se_k, ctS,
                 scale: float, se_p, qhi: int, se_v, v_aft, *, save_p: bool = False, fmt16: Optional[int] = None,
                 want_rowsum: bool = False, want_rowstat: bool = False):
    """Fused QKR attention forward (ofq_qkr_attn_fwd): scores, softmax, probability codes and P.V in ONE kernel.
    qx int8 [B*N, C], qk int8 [B*N, H*C], qvT int8 [B, C, ldv] (codes_transpose). Returns (out fp32 [B, N, C],
    qp int8 [B*H, N, ldq], P fp32 [B*H, N, ldS] | None, qp16 | None, rowsum | None[, rowstat [B*H, N, 2] with want_rowstat])."""
    _cuda(qx, qk, qvT)
    dev = qx.device
    ldq = max(round_up(N, 16), 208)
    ldS = round_up(N, 4)
    out = torch.empty((B, N, Cc), dtype=torch.float32, device=dev)
    qp = torch.empty((B * H, N, ldq), dtype=torch.int8, device=dev)
    P = torch.empty((B * H, N, ldS), dtype=torch.float32, device=dev) if save_p else None
    qp16 = torch.empty((B * H, N, ldq), dtype=_T16[fmt16], device=dev) if fmt16 is not None else None
    rowsum = torch.empty((B * H, N), dtype=torch.float32, device=dev) if want_rowsum else None
    rowstat = torch.empty((B * H, N, 2), dtype=torch.float32, device=dev) if want_rowstat else None
    nbytes = (B * N * Cc * (1.0 + H + 1.0 + 4.0) + B * H * N * (ldq * (1.0 + (2.0 if fmt16 is not None else 0.0)) + (4.0 * ldS if save_p else 0.0)))
    flops = 2.0 * B * H * N * N * (Cc + Cc // H)
    _call("qkr_attn_fwd", 1, nbytes, flops, _lib.load().ofq_qkr_attn_fwd, qx.data_ptr(), qk.data_ptr(), qvT.data_ptr(), qvT.shape[-1],
          B, N, H, Cc, se_x.data_ptr(), se_k.data_ptr(), ctS.data_ptr(), float(scale), se_p.data_ptr(), int(qhi), se_v.data_ptr(),
          v_aft.data_ptr(), qp.data_ptr(), ldq, out.data_ptr(), _ptr(P), ldS, _ptr(qp16), fmt16 if fmt16 is not None else FMT_F16,
          _ptr(rowsum), _ptr(rowstat), _st())
    if want_rowstat:
        return out, qp, P, qp16, rowsum, rowstat
    return out, qp, P, qp16, rowsum


def qkr_attn_bwd(qx, qk, a16, qv16, fmt16: int, B: int, N: int, H: int, Cc: int, se_x, se_k, ctS, scale: float, sp2, qhi: int,
                 rowstat, rowdot, sc_in, se_v, v_aft, qmax_v: int, g_s: float):
    """Fused backward of softmax + probability quantizer with the score recomputation and dP = dO v^T (ofq_qkr_attn_bwd).
    sp2 = [se_p, 1/se_p]. Returns (dS16 [B, 1, H, N, ldo], ldo, colsum [B*H, N], d_s [N], sc4 = [scale of dS16, 1/scale, 0, 0])."""
    _cuda(qx, qk, a16, qv16)
    dev = qx.device
    ldo = round_up(N, 8)
    # (zero-filled: the pitch padding is read by the GEMMs' TMA boxes)
    dS16 = torch.zeros((B, 1, H, N, ldo), dtype=_T16[fmt16], device=dev) if ldo > round_up(N, 8) else torch.empty((B, 1, H, N, ldo), dtype=_T16[fmt16], device=dev)
    colsum = torch.empty((B * H, N), dtype=torch.float32, device=dev)
    ds_part = torch.empty((B * H, N), dtype=torch.float32, device=dev)
    d_s = torch.empty(N, dtype=torch.float32, device=dev)
    sc4 = scratch_zeros(4, dev)
    nbytes = B * N * Cc * (1.0 + H + 2.0 + 2.0) + B * H * N * (2.0 * ldo + 16.0)
    flops = 2.0 * B * H * N * N * (Cc + Cc // H)
    _call("qkr_attn_bwd", 3, nbytes, flops, _lib.load().ofq_qkr_attn_bwd, qx.data_ptr(), qk.data_ptr(), a16.data_ptr(), qv16.data_ptr(),
          fmt16, B, N, H, Cc, se_x.data_ptr(), se_k.data_ptr(), ctS.data_ptr(), float(scale), sp2[0].data_ptr(), sp2[1].data_ptr(),
          int(qhi), rowstat.data_ptr(), rowdot.data_ptr(), sc_in.data_ptr(), se_v.data_ptr(), v_aft.data_ptr(), int(qmax_v), float(g_s),
          dS16.data_ptr(), ldo, colsum.data_ptr(), ds_part.data_ptr(), d_s.data_ptr(), sc4.data_ptr(), _st())
    return dS16, ldo, colsum, d_s, sc4


# ------------------------------------------------------------------------------------------------ W_qk
def wqk_compose(wq: torch.Tensor, wk: torch.Tensor, H: int) -> torch.Tensor:
    _cuda(wq, wk)
    Cc = wq.shape[1]
    hd = wq.shape[0] // H
    from . import prologue
    pro = prologue.ACTIVE if prologue.ENABLED else None
    if pro is not None and wq.is_contiguous() and wk.is_contiguous():
        job, ready = pro.get_wqk(wq, wk, H)
        if ready:
            return job.out
        out = job.out
    else:
        out = torch.empty((H * Cc, Cc), dtype=torch.float32, device=wq.device)
    _call("wqk_compose", 1, 4.0 * (2 * H * hd * Cc + H * Cc * Cc), 2.0 * H * hd * Cc * Cc, _lib.load().ofq_wqk_compose,
          wq.data_ptr(), wk.data_ptr(), H, hd, Cc, out.data_ptr(), _st())
    return out


def wqk_compose_bwd(dwqk: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, H: int, out=None):
    Cc = wq.shape[1]
    hd = wq.shape[0] // H
    dwq, dwk = out if out is not None else (torch.empty_like(wq), torch.empty_like(wk))
    _call("wqk_compose_bwd", 1, 4.0 * (4 * H * hd * Cc + 2 * H * Cc * Cc), 4.0 * H * hd * Cc * Cc,
          _lib.load().ofq_wqk_compose_bwd, dwqk.data_ptr(), wq.data_ptr(), wk.data_ptr(), H, hd, Cc, dwq.data_ptr(),
          dwk.data_ptr(), _st())
    return dwq, dwk


# ------------------------------------------------------------------------------------------------ CGA / AdamW
def cga_mask(w: torch.Tensor, bits: int, boundary_range: float) -> torch.Tensor:
    """uint8 [rows, cols]: 1 = frozen, 0 = trainable (cga.py:450-469)."""
    _cuda(w)
    assert w.dim() == 2 and w.is_contiguous() and w.dtype == torch.float32
    R, Cc = w.shape
    mask = torch.empty((R, Cc), dtype=torch.uint8, device=w.device)
    rowstat = torch.empty(R, dtype=torch.float32, device=w.device)
    mm = torch.empty(2, dtype=torch.int32, device=w.device)
    _call("cga_mask", 3, 9.0 * R * Cc, 0, _lib.load().ofq_cga_mask, w.data_ptr(), R, Cc, bits, float(boundary_range),
          mask.data_ptr(), rowstat.data_ptr(), mm.data_ptr(), _st())
    return mask


def cga_adamw_(p: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
               lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, bits: int = 0,
               boundary_range: float = 0.005, scratch=None, mask_out: Optional[torch.Tensor] = None,
               step_dev: Optional[torch.Tensor] = None) -> None:
    """In-place (masked) AdamW step on one parameter. bits == 0: plain AdamW. step_dev: device int32 step counter
    (CUDA-graph friendly) used instead of `step`."""
    _cuda(p, grad, exp_avg, exp_avg_sq)
    assert p.is_contiguous() and grad.is_contiguous() and p.dtype == torch.float32
    rows = cols = 0
    rowstat = mm = None
    if bits > 0:
        rows, cols = p.shape
        if scratch is None:
            scratch = (torch.empty(rows, dtype=torch.float32, device=p.device),
                       torch.empty(2, dtype=torch.int32, device=p.device))
        rowstat, mm = scratch
    _call("cga_adamw", 3 if bits > 0 else 1, (32.0 if bits > 0 else 28.0) * p.numel(), 0, _lib.load().ofq_cga_adamw,
          p.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), p.numel(), rows, cols, step, lr,
          beta1, beta2, eps, weight_decay, bits, float(boundary_range), _ptr(rowstat), _ptr(mm), _ptr(mask_out),
          _ptr(step_dev), _st())


def counter_increment_(counter: torch.Tensor) -> None:
    _cuda(counter)
    assert counter.dtype == torch.int32
    _call("counter", 1, 8.0, 0, _lib.load().ofq_counter_increment, counter.data_ptr(), _st())


def build_adamw_table(entries, device) -> tuple:
    """entries: list of (p, grad, exp_avg, exp_avg_sq, weight_decay). Returns (device uint8 table, n_entries, total_blocks,
    total_numel). The table holds raw pointers: it is valid as long as those tensors keep their storage (the learning
    rate is a launch argument, not part of the table)."""
    import struct
    blob = bytearray()
    first = 0
    total = 0
    for p, g, m, v, wd in entries:
        n = p.numel()
        blob += struct.pack("<QQQQqfi", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, float(wd), first)
        first += (n + 1023) // 1024
        total += n
    # pinned staging + async copy: legal inside CUDA-graph capture (the caller keeps `host` alive with the table)
    host = torch.frombuffer(blob, dtype=torch.uint8).clone().pin_memory()      # bytearray: writable, no aliasing warning
    t = host.to(device, non_blocking=True)
    t._ofq_host = host
    return t, len(entries), first, total


def adamw_multi_(table, n_entries: int, total_blocks: int, total_numel: int, step: int, lr: float, beta1: float,
                 beta2: float, eps: float, step_dev: Optional[torch.Tensor] = None) -> None:
    _call("adamw_multi", 1, 28.0 * total_numel, 0, _lib.load().ofq_adamw_multi, table.data_ptr(), n_entries,
          total_blocks, step, lr, beta1, beta2, eps, _ptr(step_dev), _st())


def build_cga_table(entries, device) -> tuple:
    """entries: list of (p [rows, cols], grad, exp_avg, exp_avg_sq, rowstat scratch [rows], kminmax scratch [2], weight_decay).
    Returns (device table, n_entries, total_blocks, total_rowblocks, total_numel) for cga_adamw_multi_."""
    import struct
    blob = bytearray()
    first = first_row = total = 0
    for p, g, m, v, rowstat, mm, wd in entries:
        rows, cols = p.shape
        blob += struct.pack("<QQQQQQiifiii", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), rowstat.data_ptr(),
                            mm.data_ptr(), rows, cols, float(wd), first, first_row, 0)
        first += (rows * cols + 1023) // 1024
        first_row += (rows + 7) // 8
        total += rows * cols
    host = torch.frombuffer(blob, dtype=torch.uint8).clone().pin_memory()
    t = host.to(device, non_blocking=True)
    t._ofq_host = host
    return t, len(entries), first, first_row, total


def cga_adamw_multi_(table, n_entries: int, total_blocks: int, total_rowblocks: int, total_numel: int, step: int, lr: float,
                     beta1: float, beta2: float, eps: float, bits: int, boundary_range: float,
                     step_dev: Optional[torch.Tensor] = None) -> None:
    """CGA-masked AdamW on every weight of the table: 3 launches (scratch init, row statistics, masked update)."""
    _call("cga_adamw", 3, 32.0 * total_numel, 0, _lib.load().ofq_cga_adamw_multi, table.data_ptr(), n_entries, total_blocks,
          total_rowblocks, step, lr, beta1, beta2, eps, bits, float(boundary_range), _ptr(step_dev), _st())


# ------------------------------------------------------------------------------------------------ KD losses
def kd_loss(z_hard: Optional[torch.Tensor], z_soft: Optional[torch.Tensor], teacher: Optional[torch.Tensor],
            target: Optional[torch.Tensor], T: float = 1.0):
    """mean_b [ CE(z_hard, target) - sum softmax(teacher / T) log_softmax(z_soft / T) ] and its gradients (ofq_kd_loss).
    Returns (loss scalar tensor, d loss / d z_hard | None, d loss / d z_soft | None); z_hard may be z_soft (single-output student:
    the summed gradient is returned as the first)."""
    ref = z_hard if z_hard is not None else z_soft
    _cuda(ref, teacher)
    B, K = ref.shape
    dev = ref.device
    same = z_hard is not None and z_soft is not None and z_hard.data_ptr() == z_soft.data_ptr()
    row = torch.empty(B, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dzh = torch.empty((B, K), dtype=torch.float32, device=dev) if z_hard is not None else None
    dzs = torch.empty((B, K), dtype=torch.float32, device=dev) if (teacher is not None and not same) else None
    for t in (z_hard, z_soft, teacher):
        assert t is None or (t.is_contiguous() and t.dtype == torch.float32 and tuple(t.shape) == (B, K))
    assert target is None or (target.dtype == torch.int64 and target.is_contiguous() and target.numel() == B)
    _call("kd_loss", 2, 4.0 * B * K * (int(z_hard is not None) * 2 + int(teacher is not None) * 3), 0, _lib.load().ofq_kd_loss,
          _ptr(z_hard), _ptr(z_soft), _ptr(teacher), _ptr(target), B, K, float(T), row.data_ptr(), loss.data_ptr(), _ptr(dzh),
          _ptr(dzs), _st())
    return loss, dzh, dzs


# ------------------------------------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x2d: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, add: Optional[torch.Tensor] = None):
    """(y, mean, rstd) of nn.LayerNorm over the rows of x2d; with `add` [rows, cols] (cols <= 512): (x2d + add, y, mean, rstd),
    the residual sum formed and normalised in one pass."""
    _cuda(x2d, gamma, beta)
    rows, cols = x2d.shape
    y = torch.empty_like(x2d)
    mean = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    if add is not None:
        _cuda(add)
        assert add.shape == x2d.shape and add.is_contiguous() and x2d.is_contiguous()
        xsum = torch.empty_like(x2d)
        _call("layernorm_fwd", 1, 16.0 * rows * cols, 0, _lib.load().ofq_layernorm_fwd_add, x2d.data_ptr(), add.data_ptr(), rows,
              cols, gamma.data_ptr(), beta.data_ptr(), float(eps), xsum.data_ptr(), y.data_ptr(), mean.data_ptr(),
              rstd.data_ptr(), _st())
        return xsum, y, mean, rstd
    _call("layernorm_fwd", 1, 8.0 * rows * cols, 0, _lib.load().ofq_layernorm_fwd, x2d.data_ptr(), rows, cols,
          gamma.data_ptr(), beta.data_ptr(), float(eps), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _st())
    return y, mean, rstd


# The latest residual-stream gradient an ofq_b200 LayerNorm backward produced, with the per-CTA maxima of |dx|: the proj / fc2
# backward that consumes exactly this tensor as dY derives its fp16 range scale from them (no absmax pass). The strong
# reference keeps the storage alive, so a matching data_ptr really is this tensor.
RESIDUAL_MAX = {"dx": None, "bmax": None, "version": -1}


def layernorm_bwd(dy2d, x2d, gamma, mean, rstd, res2d=None, want_max: bool = False):
    """dx = LayerNorm'(dy) (+ res2d: the gradient arriving over the residual connection, added in the same pass)."""
    rows, cols = x2d.shape
    lib = _lib.load()
    ws = torch.empty(lib.ofq_layernorm_bwd_workspace(rows, cols), dtype=torch.float32, device=x2d.device)
    dx = torch.empty_like(x2d)
    dgamma = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    dbeta = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    nmax = lib.ofq_layernorm_bwd_nmax(rows, cols) if want_max else 0
    bmax = torch.empty(nmax, dtype=torch.float32, device=x2d.device) if nmax > 0 else None
    _call("layernorm_bwd", 2, (12.0 + (4 if res2d is not None else 0)) * rows * cols, 0, lib.ofq_layernorm_bwd_max, dy2d.data_ptr(),
          x2d.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, cols, _ptr(res2d), dx.data_ptr(),
          dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), _ptr(bmax), _st())
    if bmax is not None:
        import weakref
        # weak reference: the maxima are only trusted while this very tensor is alive (so no other tensor can own its
        # address) and unmodified (version counter); nothing is kept alive past the backward
        RESIDUAL_MAX["dx"], RESIDUAL_MAX["bmax"], RESIDUAL_MAX["version"] = weakref.ref(dx), bmax, dx._version
    return dx, dgamma, dbeta


def residual_max_for(t: torch.Tensor):
    """Block maxima of |t| if t is (a view of the whole of) the latest LayerNorm-backward output, else None."""
    ref = RESIDUAL_MAX["dx"]
    dx = ref() if ref is not None else None
    if (dx is not None and t.data_ptr() == dx.data_ptr() and t.numel() == dx.numel() and t.is_contiguous()
            and dx._version == RESIDUAL_MAX["version"] and t._version == dx._version):
        return RESIDUAL_MAX["bmax"]
    return None
