"""GPU parity tests of the individual sm_100a kernels against the CPU oracle (same seeded inputs).
Integer codes and freeze masks: bit-exact. Floating-point outputs: tolerance written next to each check."""
import math

import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ofq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from ofq_b200 import _lib, ops
    assert _lib.load().ofq_device_ok() == 1, "B200 (sm_100) required: " + _lib.load().ofq_last_error().decode()
    return ops


def dev(t):
    return t.cuda().contiguous()


# ------------------------------------------------------------------------------------------------ StatsQ
@pytest.mark.parametrize("bits", [2, 3, 4])
def test_statsq_codes_dyadic_exact(ops, bits):
    """Weights on a dyadic grid: row sums are exact in any order => zero code mismatches allowed."""
    torch.manual_seed(0)
    w = torch.round(torch.randn(300, 384) / 8 * 1024) / 1024
    codes, colscale, sf, _, mm = ops.statsq_codes(dev(w), bits, want_minmax=True)
    rc, rsf = O.statsq_codes(w, bits)
    assert torch.equal(sf.cpu(), rsf.squeeze(1))
    assert torch.equal(codes.cpu().int(), rc)
    n = 2 ** (bits - 1)
    assert torch.equal(colscale.cpu(), rsf.squeeze(1) / (2 * n))
    k = (rc - 1) // 2
    assert mm.cpu().tolist() == [int(k.min()), int(k.max())]


@pytest.mark.parametrize("bits", [2, 4])
def test_statsq_codes_random_only_ties(ops, bits):
    """Random fp32 weights: the scale may differ by an ulp (summation order), so a code may flip only when
    the pre-round value sits within 1e-5 of a rounding boundary (SURVEY.md §7 hard part 3)."""
    torch.manual_seed(1)
    w = torch.nn.init.trunc_normal_(torch.empty(1536, 384), std=0.02)
    codes, colscale, sf, _, _ = ops.statsq_codes(dev(w), bits)
    rc, rsf = O.statsq_codes(w, bits)
    assert rel_err(sf.cpu(), rsf.squeeze(1)) < 1e-6
    assert (sf.cpu() - rsf.squeeze(1)).abs().max() <= 2 * torch.finfo(torch.float32).eps * rsf.abs().max()
    bad = codes.cpu().int() != rc
    if bad.any():
        b4, _ = O.statsq_pre_round(w, bits)
        frac = (b4[bad] - torch.floor(b4[bad]) - 0.5).abs()
        assert frac.max() < 1e-5, "a non-tie code mismatch"
    assert bad.float().mean() < 1e-5


def test_statsq_golden_and_colterm(ops):
    g = load_golden("statsq")
    for bits, exp in ((2, [1, -1, 1, -3, 1, 1, 3, -3]), (3, [1, -3, 3, -5, 1, 1, 7, -7]), (4, [3, -5, 7, -9, 1, 1, 15, -15])):
        codes, colscale, sf, _, _ = ops.statsq_codes(dev(g["kat.w"]), bits)
        assert codes.cpu().tolist() == [exp]
        assert sf.item() == pytest.approx(0.7625, rel=1e-6)
    w = g["rand.w"]
    aft = torch.randn(w.shape[1]) * 0.1
    bias = torch.randn(w.shape[0])
    codes, colscale, sf, colterm, _ = ops.statsq_codes(dev(w), 2, aft=dev(aft), bias=dev(bias))
    ref = colscale.cpu() * (codes.cpu().float() @ aft) + bias
    assert rel_err(colterm.cpu(), ref) < 1e-6      # fp32 dot of <=40 terms


# ------------------------------------------------------------------------------------------------ LSQ
def _g(hi, count):
    return 1.0 / ((hi * count) ** 0.5)


def test_lsq_effective_scale_bit_exact(ops):
    torch.manual_seed(2)
    a = torch.rand(1188) * 0.2
    a[:5] = torch.tensor([1e-7, 1e-5, 0.0, -1.0, 1.0000001e-5])
    for g in (_g(1, 128 * 384), _g(3, 25344), 0.123):
        out = ops.lsq_effective_scale(dev(a), g)
        ref = O.grad_scaled(O.floor_clip(a), g)
        assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("bit,pos", [(2, False), (2, True), (3, False), (4, False), (4, True)])
def test_lsq_codes_rows_bit_exact(ops, bit, pos):
    torch.manual_seed(3)
    B, N, Cc = 4, 198, 384
    x = torch.randn(B, N, Cc)
    if pos:
        x = torch.nn.functional.gelu(x)
    b4 = torch.randn(Cc) * 0.05
    lo, hi = O.lsq_levels(bit, pos)
    s = O.lsq_init_rows(x + b4, hi, pos) * torch.linspace(0.4, 1.6, N)
    g = _g(hi, B * Cc)
    se = ops.lsq_effective_scale(dev(s), g)
    codes = ops.lsq_quant(dev(x).view(B * N, Cc), dev(b4), se, ops.PER_ROW, N, 1, lo, hi)
    ref = O.lsq_codes_rows(x + b4, s, bit, pos)
    assert torch.equal(codes.cpu().int().view(B, N, Cc), ref)
    assert ref.min() == lo and ref.max() == hi


def test_lsq_codes_qkx_segments_and_cols(ops):
    """(B, N*H, C) view of qkx: scale per (token, head), shift of length H*C (attention.py:201-206);
    LsqQuantizer4v: scale per channel."""
    torch.manual_seed(4)
    B, N, H, Cc = 3, 50, 3, 64
    lo, hi = O.lsq_levels(2, False)
    x = torch.randn(B, N, H * Cc)
    b4 = torch.randn(H * Cc) * 0.1
    xs = (x + b4).reshape(B, N * H, Cc)
    s = O.lsq_init_rows(xs, hi, False) * torch.linspace(0.5, 1.5, N * H)
    g = _g(hi, B * Cc)
    codes = ops.lsq_quant(dev(x).view(B * N, H * Cc), dev(b4), ops.lsq_effective_scale(dev(s), g), ops.PER_ROW, N, H, lo, hi)
    ref = O.lsq_codes_rows(xs, s, 2, False).reshape(B, N, H * Cc)
    assert torch.equal(codes.cpu().int().view(B, N, H * Cc), ref)
    # per-channel
    xv = torch.randn(B, N, Cc)
    b4v = torch.randn(Cc) * 0.1
    sv = O.lsq_init_cols(xv + b4v, hi, False) * torch.linspace(0.5, 1.5, Cc)
    gv = _g(hi, B * N)
    codes = ops.lsq_quant(dev(xv).view(B * N, Cc), dev(b4v), ops.lsq_effective_scale(dev(sv), gv), ops.PER_COL, 1, 1, lo, hi)
    se = O.grad_scaled(O.floor_clip(sv), gv)
    ref = torch.round(torch.clamp((xv + b4v) / se, lo, hi)).int()
    assert torch.equal(codes.cpu().int().view(B, N, Cc), ref)


@pytest.mark.parametrize("cols,nseg,mode", [(384, 1, 0), (1536, 1, 0), (3 * 192, 3, 0), (384, 1, 1), (96, 1, 0), (6 * 384, 6, 0),
                                            (2 * 128, 2, 0), (200, 1, 1)])
def test_lsq_backward(ops, cols, nseg, mode):
    torch.manual_seed(5)
    B, N = 3, 70
    bit, pos = 2, False
    lo, hi = O.lsq_levels(bit, pos)
    seg = cols // nseg
    x = torch.randn(B, N, cols)
    b4 = (torch.randn(cols) * 0.05).requires_grad_(True)
    aft = torch.zeros(cols, requires_grad=True)
    dy = torch.randn(B, N, cols)
    xin = x.clone().requires_grad_(True)
    if mode == 0:
        xs = (xin + b4).reshape(B, N * nseg, seg)
        s = (O.lsq_init_rows(xs.detach(), hi, pos) * torch.linspace(0.4, 1.6, N * nseg)).requires_grad_(True)
        y = O.lsq_rows(xs, s, bit, pos).reshape(B, N, cols) + aft
        g = _g(hi, B * seg)
        period = N
    else:
        xs = xin + b4
        s = (O.lsq_init_cols(xs.detach(), hi, pos) * torch.linspace(0.4, 1.6, cols)).requires_grad_(True)
        y = O.lsq_cols(xs, s, bit, pos) + aft
        g = _g(hi, B * N)
        period = 1
    y.backward(dy)
    se = ops.lsq_effective_scale(dev(s.detach()), g)
    dx, ds, db4, daft = ops.lsq_bwd(dev(dy).view(B * N, cols), dev(x).view(B * N, cols), dev(b4.detach()), se, mode,
                                    period, nseg, lo, hi, g)
    mine = dx.cpu().view(B, N, cols)
    assert torch.equal(mine == 0, xin.grad == 0)                      # the STE clamp mask is exact
    assert rel_err(mine, xin.grad) < 1e-6                             # autograd computes (dy*s)/s, we pass dy
    assert rel_err(ds.cpu(), s.grad) < 1e-5                           # fp32 sums in a different order
    assert rel_err(db4.cpu(), b4.grad) < 1e-5
    assert rel_err(daft.cpu(), aft.grad) < 1e-5


@pytest.mark.parametrize("cols", [1536, 200])
def test_lsq_gelu_fused_and_fp16_copy(ops, cols):
    """fc2 of QMLP: codes = Q(GELU(h) + b4) with the activation evaluated inside the quantizer pass (cols = 1536: streaming
    kernels, cols = 200: generic kernels), the exact fp16 copy of the codes written in the same pass, and the backward that
    recomputes GELU(h) and returns the gradient w.r.t. h together with the fp16 range scale of the next GEMM operand."""
    torch.manual_seed(6)
    B, N, bit = 3, 66, 2
    lo, hi = O.lsq_levels(bit, True)
    h = (torch.randn(B, N, cols) * 1.5).requires_grad_(True)
    b4 = (torch.randn(cols) * 0.05).requires_grad_(True)
    aft = torch.zeros(cols, requires_grad=True)
    a = torch.nn.functional.gelu(h)
    xs = a + b4
    s = (O.lsq_init_rows(xs.detach(), hi, True) * torch.linspace(0.4, 1.6, N)).requires_grad_(True)
    y = O.lsq_rows(xs, s, bit, True) + aft
    dy = torch.randn(B, N, cols)
    y.backward(dy)
    g = _g(hi, B * cols)
    se = ops.lsq_effective_scale(dev(s.detach()), g)
    hd = dev(h.detach()).view(B * N, cols)
    codes, c16 = ops.lsq_quant(hd, dev(b4.detach()), se, ops.PER_ROW, N, 1, lo, hi, act=ops.ACT_GELU, fmt16=ops.FMT_F16)
    assert c16.dtype == torch.float16 and torch.equal(c16.float(), codes.float())
    # exact w.r.t. the GELU the GPU evaluates (ATen's formula); against the CPU's erf only rounding ties may differ
    a_gpu = torch.nn.functional.gelu(dev(h.detach())).cpu()
    ref_gpu = O.lsq_codes_rows(a_gpu + b4.detach(), s.detach(), bit, True)
    mine = codes.cpu().int().view(B, N, cols)
    # ... every mismatch classified: the reference's own pre-round value sits within 2e-5 * max(1, |v|) of a rounding boundary
    # k + 1/2 (a proven tie: erf / quotient evaluated in another order), nothing else may differ
    s_rows = se.cpu().view(1, N, 1)                      # effective step sizes (bit-exact against the oracle, tested above)
    for ref_codes, act_ref in ((ref_gpu, a_gpu), (O.lsq_codes_rows(xs.detach(), s.detach(), bit, True), a.detach())):
        bad = mine != ref_codes
        if bad.any():
            v = ((act_ref + b4.detach()) / s_rows)[bad]
            assert ((v - torch.floor(v) - 0.5).abs() <= 2e-5 * v.abs().clamp(min=1.0)).all(), "a code differs away from a rounding tie"
            assert (mine[bad] - ref_codes[bad]).abs().max() <= 1
    cs = torch.rand(cols) + 0.5
    dx, ds, db4, daft, sc = ops.lsq_bwd(dev(dy).view(B * N, cols), hd, dev(b4.detach()), se, ops.PER_ROW, N, 1, lo, hi, g,
                                        act=ops.ACT_GELU, next_scale=(dev(cs), se, 1.0, True))
    assert rel_err(dx.cpu().view(B, N, cols), h.grad) < 1e-5
    assert rel_err(ds.cpu(), s.grad) < 1e-4
    assert rel_err(db4.cpu(), b4.grad) < 1e-5
    assert rel_err(daft.cpu(), aft.grad) < 1e-5
    # range scale: a power of two that places max|dx| * max|cs| * max|se| inside fp16's range with head-room
    bound = dx.abs().max().item() * cs.max().item() * se.max().item()
    sc = sc.cpu()
    assert sc[0] * sc[1] == 1.0 and math.log2(sc[0].item()) == int(math.log2(sc[0].item()))
    assert 2.0 ** 14 <= bound * sc[0].item() < 2.0 ** 15


# ------------------------------------------------------------------------------------------------ GEMM engine
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (1584, 192, 192), (25344, 384, 384), (777, 200, 1536), (198, 198, 64)])
def test_gemm_i8_exact(ops, M, N, K):
    torch.manual_seed(6)
    A = torch.randint(-8, 8, (M, K), dtype=torch.int8, device="cuda")
    Bm = torch.randint(-15, 16, (N, K), dtype=torch.int8, device="cuda")
    ld = ops.round_up(N, 4)
    out = torch.full((M, ld), float("nan"), device="cuda")
    ops.gemm(ops.GEMM_I8, A, (K, 0, 0, 0), Bm, (K, 0, 0, 0), out, (ld, 0, 0), M, N, K)
    ref = A.double() @ Bm.double().T
    assert torch.equal(out[:, :N].double(), ref)                      # int32 accumulation is exact
    rs = torch.rand(198, device="cuda") + 0.5
    cs = torch.rand(N, device="cuda") + 0.5
    rt = torch.randn(198, device="cuda")
    ct = torch.randn(N, device="cuda")
    ops.gemm(ops.GEMM_I8, A, (K, 0, 0, 0), Bm, (K, 0, 0, 0), out, (ld, 0, 0), M, N, K,
             rs=ops.vec(rs, 198), cs=ops.vec(cs), rt=ops.vec(rt, 198), ct=ops.vec(ct))
    idx = torch.arange(M, device="cuda") % 198
    ref2 = ref * rs[idx].double()[:, None] * cs.double()[None] + rt[idx].double()[:, None] * ct.double()[None]
    assert rel_err(out[:, :N], ref2) < 1e-6                           # two fp32 roundings in the epilogue


@pytest.mark.parametrize("M,N,K,splits", [(384, 1536, 25344, 8), (25344, 384, 1536, 1), (500, 72, 200, 1), (64, 384, 198 * 4, 3)])
def test_gemm_bf16(ops, M, N, K, splits):
    torch.manual_seed(7)
    A = torch.randn(M, K, device="cuda").bfloat16()
    Bm = torch.randint(-3, 4, (N, K), device="cuda").bfloat16()
    out = torch.zeros((M, N), device="cuda")
    ops.gemm(ops.GEMM_BF16, A, (K, 0, 0, 0), Bm, (K, 0, 0, 0), out, (N, 0, 0), M, N, K, splits=splits, accumulate=splits > 1)
    ref = A.double() @ Bm.double().T
    assert rel_err(out, ref) < 1e-5                                   # fp32 accumulation of exact bf16 products


def test_gemm_batched_attention_layout(ops):
    """scores GEMM of QKR attention: A = x codes [b][n][c] shared over heads, B = qkx codes [b][d][h][c]."""
    torch.manual_seed(8)
    Bt, H, Nt, Cc = 3, 6, 198, 384
    qx = torch.randint(-2, 2, (Bt, Nt, Cc), dtype=torch.int8, device="cuda")
    qk = torch.randint(-2, 2, (Bt, Nt, H, Cc), dtype=torch.int8, device="cuda")
    S = torch.empty(Bt, H, Nt, 208, device="cuda")
    ops.gemm(ops.GEMM_I8, qx, (Cc, 0, 0, Nt * Cc), qk, (H * Cc, 0, Cc, Nt * H * Cc), S, (208, Nt * 208, H * Nt * 208),
             Nt, Nt, Cc, nb1=H, nb2=Bt)
    ref = torch.einsum("bnc,bdhc->bhnd", qx.double(), qk.double())
    assert torch.equal(S[..., :Nt].double(), ref)


def test_gemm_outer_k_accumulate(ops):
    """k2 outer-K loop (sum over heads) with bf16 operands."""
    torch.manual_seed(9)
    Bt, H, Nt, Cc, NP = 2, 3, 50, 64, 56
    dS = torch.zeros(Bt, H, Nt, NP, device="cuda")
    dS[..., :Nt] = torch.randn(Bt, H, Nt, Nt, device="cuda")
    kT = torch.zeros(Bt, H, Cc, NP, device="cuda")
    kT[..., :Nt] = torch.randint(-2, 2, (Bt, H, Cc, Nt), device="cuda").float()
    out = torch.empty(Bt, Nt, Cc, device="cuda")
    ops.gemm(ops.GEMM_BF16, dS.bfloat16(), (NP, Nt * NP, H * Nt * NP, 0), kT.bfloat16(), (NP, Cc * NP, H * Cc * NP, 0),
             out, (Cc, Nt * Cc, 0), Nt, Cc, Nt, k2=H, nb1=Bt)
    ref = torch.einsum("bhnd,bhcd->bnc", dS.bfloat16().double(), kT.double())
    assert rel_err(out, ref) < 1e-5
    # hi/lo planes: A laid out [b][plane][h], B indexed with k2 % H; and an operand shared by all slices
    hi = dS.bfloat16()
    lo = (dS - hi.float()).bfloat16()
    A2 = torch.stack((hi, lo), dim=1).contiguous()                       # [B, 2, H, N, NP]
    ops.gemm(ops.GEMM_BF16, A2, (NP, Nt * NP, 2 * H * Nt * NP, 0), kT.bfloat16(), (NP, Cc * NP, H * Cc * NP, 0),
             out, (Cc, Nt * Cc, 0), Nt, Cc, Nt, k2=2 * H, nb1=Bt, b_k2mod=H)
    ref2 = torch.einsum("bhnd,bhcd->bnc", dS.double(), kT.double())
    assert rel_err(out, ref2) < 2e-5
    # dual-A: hi and lo of head h are loaded in one stage (slices h and h + H), B is read once per head
    ops.gemm(ops.GEMM_BF16, A2, (NP, Nt * NP, 2 * H * Nt * NP, 0), kT.bfloat16(), (NP, Cc * NP, H * Cc * NP, 0),
             out, (Cc, Nt * Cc, 0), Nt, Cc, Nt, k2=H, nb1=Bt, a_dual_delta=H)
    assert rel_err(out, ref2) < 2e-5
    A3 = torch.stack((hi[:, 0], lo[:, 0]), dim=0).contiguous()           # [2, B, N, NP], head 0 only
    ops.gemm(ops.GEMM_BF16, A3, (NP, Bt * Nt * NP, Nt * NP, 0), kT.bfloat16()[:, 0].contiguous(), (NP, 0, Cc * NP, 0),
             out, (Cc, Nt * Cc, 0), Nt, Cc, Nt, k2=2, nb1=Bt)
    assert rel_err(out, torch.einsum("bnd,bcd->bnc", dS[:, 0].double(), kT[:, 0].double())) < 2e-5
    ops.gemm(ops.GEMM_BF16, A3, (NP, Bt * Nt * NP, Nt * NP, 0), kT.bfloat16()[:, 0].contiguous(), (NP, 0, Cc * NP, 0),
             out, (Cc, Nt * Cc, 0), Nt, Cc, Nt, k2=1, nb1=Bt, a_dual_delta=1)
    assert rel_err(out, torch.einsum("bnd,bcd->bnc", dS[:, 0].double(), kT[:, 0].double())) < 2e-5


@pytest.mark.parametrize("M,N,K,splits", [(25344, 384, 1536, 1), (1536, 384, 25344, 9), (300, 200, 500, 1), (128, 64, 64, 1)])
def test_gemm_bf16_dual_planes(ops, M, N, K, splits):
    """hi + lo planes of a real-valued operand against exact integer codes: ~16 mantissa bits, one B load per stage."""
    torch.manual_seed(18)
    Kp = ops.round_up(K, 8)
    x = torch.randn(M, K, device="cuda")
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    A = torch.zeros(2, M, Kp, device="cuda", dtype=torch.bfloat16)
    A[0, :, :K], A[1, :, :K] = hi, lo
    Bm = torch.zeros(N, Kp, device="cuda", dtype=torch.bfloat16)
    Bm[:, :K] = torch.randint(-3, 4, (N, K), device="cuda").bfloat16()
    out = torch.zeros(M, ops.round_up(N, 4), device="cuda")
    ops.gemm(ops.GEMM_BF16, A, (Kp, M * Kp, 0, 0), Bm, (Kp, 0, 0, 0), out, (out.shape[1], 0, 0), M, N, K, k2=1,
             a_dual_delta=1, splits=splits, accumulate=splits > 1)
    ref = x.double() @ Bm[:, :K].double().T
    assert rel_err(out[:, :N], ref) < 2e-5


# ------------------------------------------------------------------------------------------------ prep kernels
def test_grad_prep_and_code_conversions(ops):
    torch.manual_seed(10)
    nb, R, Cc = 3, 198, 384
    x = torch.randn(nb, R, Cc, device="cuda")
    cs = torch.rand(Cc, device="cuda") + 0.5
    rs = torch.rand(R, device="cuda") + 0.5
    u = torch.randn(Cc, device="cuda")
    o = ops.grad_prep(x, nb, R, Cc, Cc, R * Cc, cs=cs, rs=rs, rs_period=R, want_rm=True, want_t=True, want_colsum=True,
                      u=u, group=64)
    assert torch.equal(o["rm"][0], (x * cs).bfloat16())
    t_ref = torch.zeros(nb, Cc, o["r_pad"], device="cuda", dtype=torch.bfloat16)
    t_ref[..., :R] = (x * rs[None, :, None]).bfloat16().transpose(1, 2)
    assert torch.equal(o["t"][0], t_ref)
    o2 = ops.grad_prep(x, nb, R, Cc, Cc, R * Cc, cs=cs, rs=rs, rs_period=R, want_rm=True, want_t=True, planes=2)
    assert torch.equal(o2["rm"][0], o["rm"][0]) and torch.equal(o2["t"][0], o["t"][0])
    assert rel_err(o2["rm"][0].float() + o2["rm"][1].float(), x * cs) < 2e-5      # hi + lo: ~16 mantissa bits
    assert rel_err((o2["t"][0].float() + o2["t"][1].float())[..., :R], (x * rs[None, :, None]).transpose(1, 2)) < 2e-5
    assert rel_err(o["colsum"], x.sum((0, 1))) < 1e-5
    assert rel_err(o["rowdot"], (x * u).view(nb, R, Cc // 64, 64).sum(-1).transpose(1, 2)) < 1e-5
    o32 = ops.grad_prep(x, nb, R, Cc, Cc, R * Cc, u=u, group=32)
    assert rel_err(o32["rowdot"], (x * u).view(nb, R, Cc // 32, 32).sum(-1).transpose(1, 2)) < 1e-5
    rd = ops.codes_rowdot(torch.randint(-2, 2, (500, 6 * 64), dtype=torch.int8, device="cuda"), 6, u)
    assert rd.shape == (500, 6)
    codes = torch.randint(-8, 8, (nb, R, Cc), dtype=torch.int8, device="cuda")
    rd = ops.codes_rowdot(codes.view(nb * R, Cc), 6, u)
    assert rel_err(rd, (codes.float() * u).view(nb * R, 6, 64).sum(-1)) < 1e-5
    assert torch.equal(ops.codes_to_bf16(codes, nb, R, Cc, Cc, R * Cc, False), codes.bfloat16())
    tb = ops.codes_to_bf16(codes, nb, R, Cc, Cc, R * Cc, True)
    assert torch.equal(tb[..., :R], codes.bfloat16().transpose(1, 2)) and bool((tb[..., R:] == 0).all())
    ti = ops.codes_transpose(codes, nb, R, Cc, Cc, R * Cc)
    assert torch.equal(ti[..., :R], codes.transpose(1, 2)) and bool((ti[..., R:] == 0).all())


@pytest.mark.parametrize("a_mn,b_mn", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K,splits,fmt", [(1536, 384, 25344, 8, "f16"), (200, 72, 200, 1, "bf16"), (384, 2304, 1000, 2, "f16")])
def test_gemm_mn_major_operands(ops, a_mn, b_mn, M, N, K, splits, fmt):
    """MN-major (rows contiguous, K strided) 16-bit operands: a row-major [tokens][features] tensor is consumed as the
    transposed operand of dW = dY^T X without a transposed copy."""
    torch.manual_seed(14)
    dt = torch.float16 if fmt == "f16" else torch.bfloat16
    A = torch.randn(M, K, device="cuda").to(dt)
    Bm = torch.randint(-3, 4, (N, K), device="cuda").to(dt)
    a_arg, a_str = (A.t().contiguous(), (M, 0, 0, 0)) if a_mn else (A, (K, 0, 0, 0))
    b_arg, b_str = (Bm.t().contiguous(), (N, 0, 0, 0)) if b_mn else (Bm, (K, 0, 0, 0))
    out = torch.zeros((M, N), device="cuda")
    ops.gemm(ops.GEMM_F16 if fmt == "f16" else ops.GEMM_BF16, a_arg, a_str, b_arg, b_str, out, (N, 0, 0), M, N, K,
             splits=splits, accumulate=splits > 1, a_mn=a_mn, b_mn=b_mn)
    assert rel_err(out, A.double() @ Bm.double().T) < 1e-5


def test_gemm_mn_major_batched(ops):
    """Batched MN-major operands with batch strides (attention backward layout: dK_hat[d,c] = sum_n dS[n,d] x_hat[n,c])."""
    torch.manual_seed(15)
    Bt, H, N, C = 3, 2, 198, 128
    dS = torch.randn(Bt, H, N, 200, device="cuda").half()          # [b][h][n][d], pitch 200
    dS[..., N:] = 0
    xh = torch.randn(Bt, N, C, device="cuda").half()               # [b][n][c]
    out = torch.empty(Bt, N, H, C, device="cuda")                  # [b][d][h][c]
    ops.gemm(ops.GEMM_F16, dS, (200, 0, N * 200, H * N * 200), xh, (C, 0, 0, N * C), out, (H * C, C, N * H * C), N, C, N,
             nb1=H, nb2=Bt, a_mn=True, b_mn=True)
    ref = torch.einsum("bhnd,bnc->bdhc", dS[..., :N].double(), xh.double())
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize("nb,R,Cc,group", [(1, 1000, 384, 64), (3, 198, 384, 64), (4, 49, 96, 32), (1, 777, 1536, 64), (2, 50, 200, 0)])
def test_grad_prep_streaming_single_copy(ops, nb, R, Cc, group):
    """The single row-major fp16 copy of the fp16 backward (x * cs[c] * rs[r] * scale), with column sums and per-head row
    dots, produced by the streaming kernel (no transposed output requested)."""
    torch.manual_seed(21)
    x = torch.randn(nb, R, Cc, device="cuda") * 1e-4
    cs = torch.rand(Cc, device="cuda") * 0.05 + 0.01
    rs = torch.rand(R, device="cuda") + 0.5
    u = torch.randn(Cc, device="cuda") if group else None
    sc = ops.absmax_scale(x, nb, R, Cc, Cc, R * Cc, cs=cs, rs=rs, rs_period=R, product=True)
    o = ops.grad_prep(x, nb, R, Cc, Cc, R * Cc, cs=cs, rs=rs, rs_period=R, want_rm=True, want_colsum=True, u=u,
                      group=group or 64, fmt=ops.FMT_F16, scale4=sc, rm_rowscale=True)
    ref = x * cs * (sc[0] * rs)[None, :, None]
    assert rel_err(o["rm"][0].float(), ref) < 4e-4                                # fp16 rounding: 11 significant bits
    assert float((o["rm"][0].float() - ref).abs().max()) <= float(ref.abs().max()) * 2.0 ** -11
    assert rel_err(o["colsum"], x.sum((0, 1))) < 1e-5
    if group:
        rd = (x * u).view(nb, R, Cc // group, group).sum(-1).permute(0, 2, 1)
        assert rel_err(o["rowdot"], rd) < 1e-5


def test_absmax_scale_and_fp16_prep(ops):
    """Range-scaled fp16 operands (backward mode "f16"): the power-of-two scale places the operand's absmax in
    [2^14, 2^15), the fp16 planes are the correctly rounded scaled values and the GEMM un-scales in its epilogue."""
    torch.manual_seed(11)
    nb, R, Cc = 2, 198, 384
    x = torch.randn(nb, R, Cc, device="cuda") * 3e-6           # gradient-like magnitudes: far below the fp16 normal range
    x[1, 17, 5] = 2.5e-3                                       # one outlier sets the range
    cs = torch.rand(Cc, device="cuda") * 0.05 + 0.01
    rs = torch.rand(R, device="cuda") + 0.5
    sc = ops.absmax_scale(x, nb, R, Cc, Cc, R * Cc, cs=cs, rs=rs, rs_period=R)
    sc2 = ops.absmax_scale(x, nb, R, Cc, Cc, R * Cc, cs=cs, rs=rs, rs_period=R)     # the workspace counter self-resets
    assert torch.equal(sc, sc2)
    for bound, k in (((x * cs).abs().max(), 0), ((x * rs[None, :, None]).abs().max(), 2)):
        s, inv = sc[k].item(), sc[k + 1].item()
        assert s * inv == 1.0 and torch.frexp(torch.tensor(s))[0].item() == 0.5            # exact power of two
        assert 2.0 ** 14 <= bound.item() * s < 2.0 ** 15
    o = ops.grad_prep(x, nb, R, Cc, Cc, R * Cc, cs=cs, rs=rs, rs_period=R, want_rm=True, want_t=True, want_colsum=True,
                      fmt=ops.FMT_F16, scale4=sc)
    assert o["rm"].dtype == torch.float16
    assert torch.equal(o["rm"][0], (x * cs * sc[0]).half())
    assert torch.equal(o["t"][0][..., :R], (x * (rs[None, :, None] * sc[2])).half().transpose(1, 2))
    assert rel_err(o["rm"][0].float() * sc[1], x * cs) < 4e-4                     # 11 significant bits
    # bound multipliers and a ragged (N = 198, pitch 200) operand such as dL/dP_hat
    y = torch.randn(6, R, 200, device="cuda")
    y[..., R:] = float("nan")                                   # pitch padding is never read as data
    v1 = torch.rand(R, device="cuda") + 1.0
    v2 = torch.rand(R, device="cuda") + 3.0
    sc3 = ops.absmax_scale(y, 6, R, R, 200, R * 200, v1=v1, v2=v2, mult=0.25)
    amax = y[..., :R].abs().max().item()
    assert 2.0 ** 14 <= amax * v1.max().item() * 0.25 * sc3[0].item() < 2.0 ** 15
    assert 2.0 ** 14 <= amax * v2.max().item() * 0.25 * sc3[2].item() < 2.0 ** 15
    z = torch.zeros(1, 8, 64, device="cuda")
    assert ops.absmax_scale(z, 1, 8, 64, 64, 512).tolist() == [1.0, 1.0, 1.0, 1.0]           # all-zero gradient: no scaling
    codes = torch.randint(-8, 8, (nb, R, Cc), dtype=torch.int8, device="cuda")
    assert torch.equal(ops.codes_to_bf16(codes, nb, R, Cc, Cc, R * Cc, False, ops.FMT_F16), codes.half())
    tb = ops.codes_to_bf16(codes, nb, R, Cc, Cc, R * Cc, True, ops.FMT_F16)
    assert torch.equal(tb[..., :R], codes.half().transpose(1, 2)) and bool((tb[..., R:] == 0).all())


@pytest.mark.parametrize("M,N,K,splits", [(384, 1536, 25344, 8), (25344, 384, 1536, 1), (500, 72, 200, 1)])
def test_gemm_f16_scaled(ops, M, N, K, splits):
    """fp16 tensor-core GEMM with the range scale undone by a period-1 row vector in the epilogue."""
    torch.manual_seed(12)
    g = torch.randn(M, K, device="cuda") * 1e-5
    sc = ops.absmax_scale(g, 1, M, K, K, 0)
    A = (g * sc[0]).half()
    Bm = torch.randint(-3, 4, (N, K), device="cuda").half()
    out = torch.zeros((M, N), device="cuda")
    ops.gemm(ops.GEMM_F16, A, (K, 0, 0, 0), Bm, (K, 0, 0, 0), out, (N, 0, 0), M, N, K, splits=splits, accumulate=splits > 1,
             rs=ops.vec(sc[1:2], 1))
    exact = (A.double() * sc[1].double()) @ Bm.double().T
    assert rel_err(out, exact) < 1e-5                                 # fp32 accumulation of exact fp16 products
    assert rel_err(out, g.double() @ Bm.double().T) < 4e-4            # operand rounding: 11 significant bits


# ------------------------------------------------------------------------------------------------ softmax
@pytest.mark.parametrize("N,H,bit", [(198, 6, 2), (49, 3, 3), (197, 3, 4)])
def test_softmax_quant_forward(ops, N, H, bit):
    torch.manual_seed(11)
    B = 4
    lo, hi = O.lsq_levels(bit, True)
    ld = ops.round_up(N, 4)
    S = torch.randn(B * H, N, ld) * 2
    prob_ref = torch.softmax(S[..., :N], dim=-1).view(B, H, N, N)
    s = O.lsq_init_rows(prob_ref, hi, True) * torch.linspace(0.5, 1.5, N)
    g = _g(hi, B * H * N)
    se = ops.lsq_effective_scale(dev(s), g)
    P, codes, rowsum = ops.softmax_quant(dev(S), N, H, se, hi)
    Pc = P.cpu()[..., :N].view(B, H, N, N)
    assert rel_err(Pc, prob_ref) < 1e-6                                 # expf / sum order differ by ulps
    # codes are exact for the probabilities the kernel itself produced ...
    ref_codes = O.lsq_codes_rows(Pc, s, bit, True)
    mine = codes.cpu()[..., :N].int().view(B, H, N, N)
    assert torch.equal(mine, ref_codes)
    assert bool((codes.cpu()[..., N:] == 0).all())
    # ... and differ from the codes of torch's own softmax only at rounding ties (SURVEY.md §7 hard part 4)
    ref2 = O.lsq_codes_rows(prob_ref, s, bit, True)
    bad = mine != ref2
    if bad.any():
        se_c = se.cpu().view(1, 1, N, 1)
        v = (prob_ref / se_c)[bad]
        assert ((v - torch.floor(v) - 0.5).abs() <= 2e-5 * v.abs().clamp(min=1.0)).all()     # proven ties only
        assert (mine[bad] - ref2[bad]).abs().max() <= 1
    assert rel_err(rowsum.cpu().view(B, H, N), ref_codes.float().sum(-1) * se.cpu().view(1, 1, N)) < 1e-6


def test_softmax_quant_bias_mask(ops):
    torch.manual_seed(12)
    B, nW, H, N = 8, 4, 3, 49
    S = torch.randn(B * H, N, 52)
    bias = torch.randn(H, N, N)
    mask = (torch.rand(nW, N, N) > 0.7).float() * -100.0
    se = torch.full((N,), 0.02)
    P, codes, _ = ops.softmax_quant(dev(S), N, H, dev(se), 3, bias=dev(bias), mask=dev(mask), nW=nW)
    logits = S[..., :N].view(B, H, N, N) + bias[None] + mask[torch.arange(B) % nW][:, None]
    assert rel_err(P.cpu()[..., :N].view(B, H, N, N), torch.softmax(logits, -1)) < 1e-6


def test_softmax_quant_backward(ops):
    torch.manual_seed(13)
    B, H, N, bit = 3, 2, 70, 2
    lo, hi = O.lsq_levels(bit, True)
    ld = 72
    S = (torch.randn(B, H, N, N) * 2).requires_grad_(True)
    alpha = 0.125
    prob = torch.softmax(S * alpha, dim=-1)
    s = (O.lsq_init_rows(prob.detach(), hi, True) * torch.linspace(0.5, 1.5, N)).requires_grad_(True)
    pq = O.lsq_rows(prob, s, bit, True)
    dPq = torch.randn(B, H, N, N)
    pq.backward(dPq)
    g = _g(hi, B * H * N)
    se = ops.lsq_effective_scale(dev(s.detach()), g)
    Sp = torch.zeros(B * H, N, ld)
    Sp[..., :N] = (S.detach() * alpha).view(B * H, N, N)
    P, codes, _ = ops.softmax_quant(dev(Sp), N, H, se, hi)
    dP = torch.zeros(B * H, N, ld)
    dP[..., :N] = dPq.view(B * H, N, N)
    ca = torch.rand(H, N) + 0.5
    rb = torch.rand(N) + 0.5
    out_a, out_bt, ldo, colsum, d_s, ds32 = ops.softmax_quant_bwd(dev(dP), P, N, H, se, hi, alpha, g, dev(ca), True, dev(rb),
                                                                  want_ds32=True)
    dS = ds32.cpu()[..., :N].view(B, H, N, N) * alpha
    assert rel_err(dS, S.grad) < 1e-5
    assert rel_err(d_s.cpu(), s.grad) < 1e-4
    assert rel_err(colsum.cpu().view(B, H, N), dS.sum(2)) < 1e-4
    ra = (dS * ca.view(1, H, 1, N)).bfloat16().float()
    assert rel_err(out_a.cpu()[..., :N].float().view(B, H, N, N), ra) < 1e-3      # bf16 output rounding
    rbt = (dS * rb.view(1, 1, N, 1)).bfloat16().float().transpose(2, 3)
    assert rel_err(out_bt.cpu()[..., :N].float().view(B, H, N, N), rbt) < 1e-3
    a2, bt2, _, _, _, _ = ops.softmax_quant_bwd(dev(dP), P, N, H, se, hi, alpha, g, dev(ca), True, dev(rb), planes=2)
    assert a2.shape == (B, 2, H, N, ldo)
    assert rel_err((a2[:, 0].float() + a2[:, 1].float()).cpu()[..., :N], dS * ca.view(1, H, 1, N)) < 2e-5
    assert rel_err((bt2[:, 0].float() + bt2[:, 1].float()).cpu()[..., :N], (dS * rb.view(1, 1, N, 1)).transpose(2, 3)) < 2e-5


@pytest.mark.parametrize("N,H,B", [(198, 6, 3), (197, 3, 2), (49, 3, 5)])
def test_softmax_quant_vectorised_single_f16(ops, N, H, B):
    """The 16-byte-aligned fast paths (forward, and the single-copy fp16 backward the DeiT step uses): pitch padding of the
    inputs is undefined (NaN here) and must never be read into a result; padding of the outputs is zero."""
    torch.manual_seed(14)
    bit = 2
    lo, hi = O.lsq_levels(bit, True)
    ld = ops.round_up(N, 4)
    S = (torch.randn(B, H, N, N) * 2).requires_grad_(True)
    alpha = 0.125
    prob = torch.softmax(S * alpha, dim=-1)
    s = (O.lsq_init_rows(prob.detach(), hi, True) * torch.linspace(0.5, 1.5, N)).requires_grad_(True)
    pq = O.lsq_rows(prob, s, bit, True)
    dPq = torch.randn(B, H, N, N)
    pq.backward(dPq)
    g = _g(hi, B * H * N)
    se = ops.lsq_effective_scale(dev(s.detach()), g)
    Sp = torch.full((B * H, N, ld), float("nan"))
    Sp[..., :N] = (S.detach() * alpha).view(B * H, N, N)
    P, codes, rowsum = ops.softmax_quant(dev(Sp), N, H, se, hi)
    Pc = P.cpu()[..., :N].view(B, H, N, N)
    assert rel_err(Pc, prob.detach()) < 1e-6
    ref_codes = O.lsq_codes_rows(Pc, s.detach(), bit, True)
    assert torch.equal(codes.cpu()[..., :N].int().view(B, H, N, N), ref_codes)      # exact for the stored probabilities
    assert bool((codes.cpu()[..., N:] == 0).all())
    assert rel_err(rowsum.cpu().view(B, H, N), ref_codes.float().sum(-1) * se.cpu().view(1, 1, N)) < 1e-6
    dP = torch.full((B * H, N, ld), float("nan"))
    dP[..., :N] = dPq.view(B * H, N, N)
    ca = torch.rand(H, N) + 0.5
    rb = torch.rand(N) + 0.5
    sc = torch.tensor([64.0, 1.0 / 64.0, 1.0, 1.0])
    out_a, out_bt, ldo, colsum, d_s, ds32 = ops.softmax_quant_bwd(dev(dP), P, N, H, se, hi, alpha, g, dev(ca), True, dev(rb),
                                                                  want_ds32=True, fmt=ops.FMT_F16, scale4=dev(sc), single=True)
    assert out_bt is None and out_a.dtype == torch.float16
    dS = ds32.cpu()[..., :N].view(B, H, N, N) * alpha
    assert rel_err(dS, S.grad) < 1e-5
    assert rel_err(d_s.cpu(), s.grad) < 1e-4
    assert rel_err(colsum.cpu().view(B, H, N), dS.sum(2)) < 1e-4
    ref = (dS * ca.view(1, H, 1, N) * rb.view(1, 1, N, 1) * 64.0).half().float()
    oa = out_a.cpu().view(B, H, N, ldo)
    assert rel_err(oa[..., :N].float(), ref) < 1e-3                                  # fp16 output rounding
    assert bool((oa[..., N:] == 0).all())


# ------------------------------------------------------------------------------------------------ W_qk
def test_wqk_compose(ops):
    torch.manual_seed(14)
    H, hd, Cc = 6, 64, 384
    wq = torch.randn(H * hd, Cc) * 0.02
    wk = torch.randn(H * hd, Cc) * 0.02
    out = ops.wqk_compose(dev(wq), dev(wk), H)
    ref = O.wqk_compose(wq.double(), wk.double(), H)
    assert rel_err(out.cpu(), ref) < 1e-6                               # fp32 FMA chain over 64 terms
    d = torch.randn(H * Cc, Cc)
    wq_, wk_ = wq.clone().double().requires_grad_(True), wk.clone().double().requires_grad_(True)
    O.wqk_compose(wq_, wk_, H).backward(d.double())
    dwq, dwk = ops.wqk_compose_bwd(dev(d), dev(wq), dev(wk), H)
    assert rel_err(dwq.cpu(), wq_.grad) < 1e-6 and rel_err(dwk.cpu(), wk_.grad) < 1e-6


# ------------------------------------------------------------------------------------------------ quantizing GEMM epilogue
@pytest.mark.parametrize("Bt,N,C,H,bits", [(16, 198, 384, 6, 2), (8, 198, 192, 3, 4), (37, 49, 96, 3, 3), (3, 198, 384, 6, 2),
                                          (1, 50, 128, 2, 2)])
def test_gemm_lsq_epilogue_is_gemm_then_quantizer(ops, Bt, N, C, H, bits):
    """ofq_gemm_lsq (the qkx GEMM with the LSQ quantizer as its epilogue, attention.py:200-207) against ofq_gemm followed by
    ofq_lsq_quant_ex on the fp32 product: codes and their 16-bit copy bit-identical, the logits' column term equal up to fp32
    summation order, the fp16 residual plane == what the backward formula makes from the fp32 product, and the LSQ backward
    fed with that plane == the backward fed with the product (same straight-through mask, same gradients)."""
    torch.manual_seed(100 + C + Bt)
    M = Bt * N
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    qx = torch.randint(lo, hi + 1, (M, C), dtype=torch.int8, device="cuda")
    wc = (torch.randint(lo, hi + 1, (H * C, C), device="cuda") * 2 + 1).to(torch.int8)            # odd StatsQ codes
    se_x = torch.rand(N, device="cuda") * 0.2 + 0.1
    cs = torch.rand(H * C, device="cuda") * 0.02 + 0.01
    ct = torch.randn(H * C, device="cuda") * 0.05
    b4 = torch.randn(H * C, device="cuda") * 0.05
    u = torch.randn(H * C, device="cuda") * 0.1
    alpha = torch.rand(N * H, device="cuda") * 0.5 + 0.2
    s2 = ops.lsq_effective_scale(alpha, 0.01, recip=True)
    vec = ops.vec
    # reference path: GEMM -> fp32 -> quantizer pass
    y = torch.empty((M, H * C), dtype=torch.float32, device="cuda")
    ops.gemm(ops.GEMM_I8, qx, (C, 0, 0, 0), wc, (C, 0, 0, 0), y, (H * C, 0, 0), M, H * C, C, rs=vec(se_x, N), cs=vec(cs), ct=vec(ct))
    fused_dot_ok = C % 128 == 0
    codes_r, c16_r, dot_r = ops.lsq_quant(y, b4, s2[0], ops.PER_ROW, N, H, lo, hi, fmt16=ops.FMT_F16, dot_u=u)
    codes, c16, res, dot = ops.gemm_lsq(qx, wc, M, H * C, C, b4, s2, N, H, lo, hi, rs=vec(se_x, N), cs=vec(cs), ct=vec(ct),
                                        fmt16=ops.FMT_F16, want_res=True, dot_u=u)
    assert torch.equal(codes, codes_r)
    assert torch.equal(c16, c16_r) and torch.equal(c16.float(), codes.float())
    assert rel_err(dot, dot_r) < 1e-5
    # codes only (the eval path) and bf16 copy
    codes_e, c16_e, res_e, dot_e = ops.gemm_lsq(qx, wc, M, H * C, C, b4, s2, N, H, lo, hi, rs=vec(se_x, N), cs=vec(cs), ct=vec(ct), dot_u=u)
    assert torch.equal(codes_e, codes) and c16_e is None and res_e is None and torch.equal(dot_e, dot)
    c16_b = ops.gemm_lsq(qx, wc, M, H * C, C, b4, s2, N, H, lo, hi, rs=vec(se_x, N), cs=vec(cs), ct=vec(ct), fmt16=ops.FMT_BF16)[1]
    assert c16_b.dtype == torch.bfloat16 and torch.equal(c16_b.float(), codes.float())
    # residual plane: q - v inside the clamp range, -2 / +2 outside, v = (y + b4) * (1 / s) as the backward evaluates it
    inv = s2[1].view(N, H).repeat(Bt, 1).repeat_interleave(C, dim=1)
    v = (y + b4) * inv
    inside = (v >= lo) & (v <= hi)
    expect = torch.where(inside, codes.float() - v, torch.where(v < lo, torch.full_like(v, -2.0), torch.full_like(v, 2.0)))
    assert torch.equal(res, expect.half())
    assert 0.005 < (~inside).float().mean().item() < 0.995                   # both regions are exercised
    # backward through the quantizer from the plane == from the fp32 product
    if C % 128 == 0:
        dy = torch.randn(M, H * C, device="cuda")
        sc4 = torch.tensor([64.0, 1 / 64.0, 0.0, 0.0], device="cuda")
        o16 = (ops.FMT_F16, cs, se_x, N, sc4)
        _, ds_a, db4_a, _, a16_a = ops.lsq_bwd(dy, y, b4, s2[0], ops.PER_ROW, N, H, lo, hi, 0.01, want_aft=False, out16=o16, want_dx=False)
        _, ds_b, db4_b, _, a16_b = ops.lsq_bwd(dy, res, b4, s2[0], ops.PER_ROW, N, H, lo, hi, 0.01, want_aft=False, out16=o16,
                                               want_dx=False, act=ops.ACT_RES16)
        assert torch.equal(a16_a, a16_b) and torch.equal(db4_a, db4_b)
        # step-size gradient: sum of dy * (q - v) with q - v rounded to fp16 (|error| <= 2^-13 per element, no bias)
        assert rel_err(ds_b, ds_a) < 3e-4


@pytest.mark.parametrize("Bt,N,K,Nout,bits", [(16, 198, 384, 384, 2), (8, 198, 384, 1536, 2), (5, 198, 192, 768, 4), (3, 50, 128, 256, 3),
                                             (1, 40, 64, 64, 2)])
def test_gemm_dx_lsq_epilogue_is_gemm_then_lsq_backward(ops, Bt, N, K, Nout, bits):
    """ofq_gemm_dx_lsq (the dX GEMM of a quantized linear layer with the backward of its input quantizer as the epilogue) against
    ofq_gemm followed by ofq_lsq_bwd on the fp32 dX_hat: dx bit-identical (same roundings, same straight-through mask), the
    step-size / shift gradients equal up to fp32 summation order."""
    torch.manual_seed(200 + K + Nout)
    M = Bt * N
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    a16 = (torch.randn(M, Nout, device="cuda") * 300).half()                   # range-scaled gradient operand
    wc8 = (torch.randint(lo, hi + 1, (Nout, K), device="cuda") * 2 + 1).to(torch.int8)
    wc = wc8.half()                                                            # exact fp16 copy of StatsQ codes (MN-major B)
    x = torch.randn(M, K, device="cuda") * 1.2
    b4 = torch.randn(K, device="cuda") * 0.05
    alpha = torch.rand(N, device="cuda") * 0.5 + 0.3
    g = 0.01
    s2 = ops.lsq_effective_scale(alpha, g, recip=True)
    sc = torch.tensor([512.0, 1 / 512.0, 0.0, 0.0], device="cuda")
    vec = ops.vec
    dxhat = torch.empty((M, K), dtype=torch.float32, device="cuda")
    ops.gemm(ops.GEMM_F16, a16, (Nout, 0, 0, 0), wc, (K, 0, 0, 0), dxhat, (K, 0, 0), M, K, Nout, b_mn=True, rs=vec(s2[1], N), cs=vec(sc[1:2], 1))
    dx_r, ds_r, db4_r, daft_r = ops.lsq_bwd(dxhat, x, b4, s2[0], ops.PER_ROW, N, 1, lo, hi, g)
    amax = torch.zeros(1, device="cuda")
    dx, ds, db4, daft = ops.gemm_dx_lsq(ops.GEMM_F16, a16, (Nout, 0, 0, 0), wc, (K, 0, 0, 0), M, K, Nout, rs=vec(s2[1], N), cs=vec(sc[1:2], 1),
                                        x2d=x, b4=b4, period=N, qlo=lo, qhi=hi, g=g, w_codes=wc8,
                                        dy_colsum=(a16.float() * s2[1].repeat(Bt).view(-1, 1) * sc[1]).sum(0), b_mn=True, amax=amax)
    assert torch.equal(dx, dx_r)
    assert amax.item() == dx_r.abs().max().item()                               # max |dx| for the consumer's fp16 range scale
    assert 0.002 < (dx == 0).float().mean().item() < 0.98                       # both sides of the mask are exercised
    assert rel_err(ds, ds_r) < 1e-5 and rel_err(db4, db4_r) < 1e-5 and rel_err(daft, daft_r) < 1e-5


# ------------------------------------------------------------------------------------------------ CGA
@pytest.mark.parametrize("bits,br", [(2, 0.005), (3, 0.005), (4, 0.05)])
def test_cga_mask_bit_exact(ops, bits, br):
    g = load_golden("cga")
    m = ops.cga_mask(dev(g["w"]), bits, br)
    key = f"mask.b{bits}.br{br}"
    if key in g:
        assert torch.equal(m.cpu().float(), g[key])                    # reference's own mask
    torch.manual_seed(15)
    w = torch.round(torch.nn.init.trunc_normal_(torch.empty(384, 1536), std=0.02) * 2 ** 16) / 2 ** 16  # dyadic
    m = ops.cga_mask(dev(w), bits, br)
    assert torch.equal(m.cpu().float(), O.cga_freeze_mask(w, bits, br))


def test_cga_masked_adamw(ops):
    g = load_golden("cga")
    for step in range(3):
        # every step starts from the reference's own previous state so that one-ulp AdamW differences do not
        # accumulate into a different mask
        if step == 0:
            w = dev(g["w"].clone())
            m = torch.zeros_like(w)
            v = torch.zeros_like(w)
        else:
            w = dev(g[f"step{step - 1}.w"].clone())
            m = dev(g[f"step{step - 1}.exp_avg"].clone())
            v = dev(g[f"step{step - 1}.exp_avg_sq"].clone())
        mask = torch.empty(w.shape, dtype=torch.uint8, device="cuda")
        before = w.clone()
        ops.cga_adamw_(w, dev(g[f"step{step}.grad"]), m, v, step + 1, 1e-3, 0.9, 0.999, 1e-8, 0.05, bits=2,
                       boundary_range=0.05, mask_out=mask)
        assert torch.equal(mask.cpu().float(), g[f"step{step}.mask"])
        frozen = mask.bool()
        assert torch.equal(w[frozen], before[frozen])                   # frozen weights untouched bit-for-bit
        assert torch.equal(w.cpu()[frozen.cpu()], g[f"step{step}.w"][frozen.cpu()])
        assert rel_err(w.cpu(), g[f"step{step}.w"]) < 1e-6
        assert rel_err(m.cpu(), g[f"step{step}.exp_avg"]) < 1e-6
        assert rel_err(v.cpu(), g[f"step{step}.exp_avg_sq"]) < 1e-6
    # unmasked path == torch.optim.AdamW
    p = torch.randn(1000, device="cuda")
    ref = torch.nn.Parameter(p.clone().cpu())
    opt = torch.optim.AdamW([ref], lr=3e-3, weight_decay=0.01)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(4):
        gr = torch.randn(1000)
        ref.grad = gr.clone()
        opt.step()
        ops.cga_adamw_(p, dev(gr), m, v, step + 1, 3e-3, 0.9, 0.999, 1e-8, 0.01)
    assert rel_err(p.cpu(), ref.detach()) < 1e-6


def test_adamw_multi_tensor(ops):
    """One launch over a pointer table == torch.optim.AdamW on every tensor (two weight-decay groups)."""
    from ofq_b200.cga import CGAAdamW
    torch.manual_seed(16)
    shapes = [(384, 384), (1536,), (7,), (1000, 384), (1,), (198,), (3, 5, 7)]
    mine = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().cpu().clone()) for p in mine]
    groups = lambda ps: [{"params": [p for p in ps if p.ndim <= 1], "weight_decay": 0.0},
                         {"params": [p for p in ps if p.ndim > 1], "weight_decay": 0.05}]
    opt = CGAAdamW(groups(mine), lr=2e-3)
    ropt = torch.optim.AdamW(groups(ref), lr=2e-3)
    for step in range(5):
        for p, r in zip(mine, ref):
            g = torch.randn(r.shape)
            r.grad = g.clone()
            p.grad = g.cuda()
        opt.step()
        ropt.step()
    for p, r in zip(mine, ref):
        assert rel_err(p.detach().cpu(), r.detach()) < 1e-6


def test_cga_adamw_multi_tensor_is_the_per_tensor_kernel(ops):
    """CGAAdamW's three-launch multi-tensor masked step == ofq_cga_adamw on each weight, bit for bit (weights, both
    moments), over steps with a changing learning rate (lr is a launch argument: the pointer tables must survive it)."""
    from ofq_b200.cga import CGAAdamW
    torch.manual_seed(18)
    shapes = [(384, 384), (1536, 384), (384, 1536), (1152, 384), (96, 100), (10, 7), (384,), (1000, 384)]
    ps = [torch.nn.Parameter(torch.nn.init.trunc_normal_(torch.empty(s), std=0.02).cuda()) for s in shapes]
    masked = [p for p in ps[:6]]
    groups = [{"params": [p for p in ps if p.ndim <= 1], "weight_decay": 0.0},
              {"params": [p for p in ps if p.ndim > 1], "weight_decay": 0.05}]
    opt = CGAAdamW(groups, lr=1e-3, masked=masked, wq_bitw=2, boundary_range=0.05)
    ref = [p.detach().clone() for p in ps]
    rm = [torch.zeros_like(p) for p in ps]
    rv = [torch.zeros_like(p) for p in ps]
    tables = None
    for step in range(4):
        lr = 1e-3 * (1.0 - 0.2 * step)
        for g in opt.param_groups:
            g["lr"] = lr
        grads = [torch.randn_like(p) * 1e-3 for p in ps]
        for p, g in zip(ps, grads):
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)                                           # persistent gradient buffers, as under the captured step
        before = opt.launches
        opt.step()
        assert opt.launches - before == 2 + 3                             # one plain launch per group + the masked trio
        if tables is None:
            tables = {k: v[1].data_ptr() for k, v in opt._tables.items()}
        else:
            assert tables == {k: v[1].data_ptr() for k, v in opt._tables.items()}      # not rebuilt when lr changed
        sd = torch.full((1,), step + 1, dtype=torch.int32, device="cuda")     # both sides form the bias corrections on the device
        for i, (r, g) in enumerate(zip(ref, grads)):
            is_masked = i < 6
            ops.cga_adamw_(r, g, rm[i], rv[i], step + 1, lr, 0.9, 0.999, 1e-8, 0.05 if r.ndim > 1 else 0.0,
                           bits=2 if is_masked else 0, boundary_range=0.05, step_dev=sd)
        for i, (p, r) in enumerate(zip(ps, ref)):
            assert torch.equal(p.detach(), r), (step, i)
            assert torch.equal(opt.state[p]["exp_avg"], rm[i]), (step, i)
            assert torch.equal(opt.state[p]["exp_avg_sq"], rv[i]), (step, i)


@pytest.mark.parametrize("rows,cols", [(25344, 384), (1000, 1536), (396, 64), (77, 96)])
def test_layernorm_fwd_bwd(ops, rows, cols):
    """Host-glue LayerNorm kernels vs torch.nn.functional.layer_norm (fp64 reference)."""
    torch.manual_seed(17)
    x = torch.randn(rows, cols, device="cuda") * 2 + 0.5
    g = torch.rand(cols, device="cuda") + 0.5
    b = torch.randn(cols, device="cuda")
    dy = torch.randn(rows, cols, device="cuda")
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6)
    xd, gd, bd = x.double().requires_grad_(True), g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xd, (cols,), gd, bd, 1e-6)
    yr.backward(dy.double())
    assert rel_err(y, yr.detach()) < 1e-6
    dx, dg, db = ops.layernorm_bwd(dy, x, g, mean, rstd)
    assert rel_err(dx, xd.grad) < 1e-5 and rel_err(dg, gd.grad) < 1e-5 and rel_err(db, bd.grad) < 1e-5
    # gradient of the residual connection around the LayerNorm added in the same pass (host/layers.py forward_res)
    res = torch.randn(rows, cols, device="cuda")
    dx2, dg2, db2 = ops.layernorm_bwd(dy, x, g, mean, rstd, res)
    assert rel_err(dx2, xd.grad + res.double()) < 1e-5 and rel_err(dg2, dg) < 1e-5 and rel_err(db2, db) < 1e-5


def test_layernorm_residual_module(ops):
    """x + f(LayerNorm(x)) through LayerNorm.forward_res equals the plain composition (value and gradients)."""
    from ofq_b200.host.layers import LayerNorm
    torch.manual_seed(18)
    ln = LayerNorm(384, eps=1e-6).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_()
    x0 = torch.randn(4, 50, 384, device="cuda")
    w = torch.randn(384, 384, device="cuda") * 0.05
    outs = []
    for fused in (True, False):
        x = x0.clone().requires_grad_(True)
        h = x * 1.5
        ln.zero_grad()
        if fused:
            r, y = ln.forward_res(h)
        else:
            r, y = h, ln(h)
        out = r + torch.tanh(y @ w)
        out.square().sum().backward()
        outs.append((out.detach(), x.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone()))
    for a, b in zip(*outs):
        assert rel_err(a, b) < 1e-6


@pytest.mark.parametrize("rows,cols", [(25344, 384), (396, 192), (77, 96), (130, 512)])
def test_layernorm_fwd_add(ops, rows, cols):
    """x + a and LayerNorm(x + a) in one pass: bit-identical to the separate add + ofq_layernorm_fwd."""
    torch.manual_seed(19)
    x = torch.randn(rows, cols, device="cuda") * 2 + 0.5
    a = torch.randn(rows, cols, device="cuda")
    g = torch.rand(cols, device="cuda") + 0.5
    b = torch.randn(cols, device="cuda")
    xs, y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6, add=a)
    y0, mean0, rstd0 = ops.layernorm_fwd(x + a, g, b, 1e-6)
    assert torch.equal(xs, x + a) and torch.equal(y, y0) and torch.equal(mean, mean0) and torch.equal(rstd, rstd0)


def test_layernorm_add_module(ops):
    """LayerNorm.forward_res_add(x, a) equals x + a followed by forward_res (values bit-identical, gradients equal)."""
    from ofq_b200.host.layers import LayerNorm
    torch.manual_seed(20)
    ln = LayerNorm(384, eps=1e-6).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_()
    x0 = torch.randn(4, 50, 384, device="cuda")
    a0 = torch.randn(4, 50, 384, device="cuda")
    w = torch.randn(384, 384, device="cuda") * 0.05
    outs = []
    for fused in (True, False):
        x = x0.clone().requires_grad_(True)
        a = a0.clone().requires_grad_(True)
        ln.zero_grad()
        if fused:
            r, y = ln.forward_res_add(x * 1.5, a * 0.5)
        else:
            r, y = ln.forward_res(x * 1.5 + a * 0.5)
        out = r + torch.tanh(y @ w)
        out.square().sum().backward()
        outs.append((out.detach(), x.grad.clone(), a.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    for u, v in zip(*outs):
        assert rel_err(u, v) < 1e-6


def test_gemm_pair_kernel_forced(ops):
    """The CTA-pair (cta_group::2) kernel is chosen by a shape heuristic; OFQ_GEMM_PAIR=2 forces it wherever it is legal
    (M > 128, no dual-A), so the whole GEMM test matrix (int8 exact, fp16 / bf16, MN-major, batched, split-K, outer-K) runs
    on it in a child process (the switch is read once per process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, OFQ_GEMM_PAIR="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "gemm and not forced", "-p", "no:cacheprovider"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_gemm_16_epilogue_warp_variant_forced(ops):
    """gemm_tc16_kernel (16 epilogue warps, 16-column TMEM pieces, 96 registers) is a measured-and-not-adopted variant of the
    single-CTA GEMM (no faster than 8 warps on B200: the store-heavy tiles are bound by shared-memory / L2 bandwidth, not by
    epilogue latency); OFQ_GEMM_EPI16=1 selects it, and the whole GEMM test matrix must still hold on it."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, OFQ_GEMM_EPI16="1", OFQ_GEMM_PAIR="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "gemm and not forced", "-p", "no:cacheprovider"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
