"""The fused QKR attention forward (ofq_qkr_attn_fwd: scores -> softmax -> probability codes -> P.V in ONE tcgen05 kernel,
logits never in HBM; reference attention.py:210-219) against the three-kernel path it replaces (int8 score GEMM ->
ofq_softmax_quant -> int8 P.V GEMM), which the golden / oracle tests pin to the reference: every output must be
BIT-IDENTICAL (the fused kernel reproduces the same fp32 operation order), at the DeiT-S / DeiT-T shapes and at ragged ones."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from ofq_b200 import _lib, ops
    from ofq_b200.quantization import functional as Fn
    assert _lib.load().ofq_device_ok() == 1
    return ops, Fn


def _unfused(ops, Fn, qx, qk, qv, B, N, H, C, se_x, se_k, ctS, scale, se_p, qhi, se_v, v_aft, fmt16):
    from ofq_b200.ops import GEMM_I8, round_up, vec
    se_k_hn = se_k.view(N, H).t().contiguous()
    cs_S = se_k_hn * scale
    ct_S = (ctS.view(B, N, H).permute(0, 2, 1) * cs_S.unsqueeze(0)).contiguous()
    ldS = round_up(N, 4)
    S = torch.empty((B * H, N, ldS), dtype=torch.float32, device=qx.device)
    ops.gemm(GEMM_I8, qx, (C, 0, 0, N * C), qk, (H * C, 0, C, N * H * C), S, (ldS, N * ldS, H * N * ldS),
             N, N, C, nb1=H, nb2=B, rs=vec(se_x, N), cs=vec(cs_S, 0, N), ct=vec(ct_S, 0, N, H * N))
    P, qp, rowsum, qp16 = ops.softmax_quant(S, N, H, se_p, qhi, save_p=True, fmt16=fmt16)
    out = Fn._pv_forward(qp, qp.shape[-1], rowsum, qv, se_p, se_v, v_aft, B, N, H, C)
    return out, qp, P, qp16, rowsum


@pytest.mark.parametrize("B,N,H,bits", [(3, 198, 6, 2), (5, 198, 3, 2), (2, 198, 6, 4), (2, 40, 2, 3), (1, 129, 1, 2),
                                        (150, 198, 6, 2), (7, 208, 4, 2), (2, 128, 2, 2), (3, 17, 5, 4)])
def test_fused_forward_bit_identical_to_three_kernel_path(mods, B, N, H, bits):
    ops, Fn = mods
    C = 64 * H
    torch.manual_seed(B * 1000 + N + H)
    dev = "cuda"
    lo, hi = -(2 ** (bits - 1)), 2 ** (bits - 1) - 1
    qhi = 2 ** bits - 1
    qx = torch.randint(lo, hi + 1, (B * N, C), dtype=torch.int8, device=dev)
    qk = torch.randint(lo, hi + 1, (B * N, H * C), dtype=torch.int8, device=dev)
    qv = torch.randint(lo, hi + 1, (B * N, C), dtype=torch.int8, device=dev)
    se_x = torch.rand(N, device=dev) * 0.5 + 0.5
    se_k = (torch.rand(N * H, device=dev) * 0.5 + 0.5) * (2.0 / (C ** 0.5) / max(1, 2 ** (bits - 2)) ** 2)
    ctS = torch.randn(B * N, H, device=dev)
    se_p = torch.rand(N, device=dev) * (0.5 / qhi) + 0.2 / qhi
    se_v = torch.rand(C, device=dev) * 0.1 + 0.05
    v_aft = torch.randn(C, device=dev) * 0.02
    scale = 64 ** -0.5
    for fmt16 in (ops.FMT_F16, ops.FMT_BF16):
        ref = _unfused(ops, Fn, qx, qk, qv, B, N, H, C, se_x, se_k, ctS, scale, se_p, qhi, se_v, v_aft, fmt16)
        qvT = ops.codes_transpose(qv, B, N, C, C, N * C)
        out, qp, P, qp16, rowsum = ops.qkr_attn_fwd(qx, qk, qvT, B, N, H, C, se_x, se_k, ctS, scale, se_p, qhi, se_v, v_aft,
                                                    save_p=True, fmt16=fmt16, want_rowsum=True)
        torch.cuda.synchronize()
        o_r, qp_r, P_r, qp16_r, rs_r = ref
        w = qp_r.shape[-1]
        assert int(qp.max()) <= qhi and int(qp.min()) >= 0
        assert torch.equal(qp[..., :w], qp_r), f"codes differ in {int((qp[..., :w] != qp_r).sum())} places"
        assert bool((qp[..., N:] == 0).all())
        assert torch.equal(P[..., :N], P_r[..., :N])
        assert torch.equal(qp16[..., :w], qp16_r)
        assert torch.equal(rowsum, rs_r)
        assert torch.equal(out, o_r), f"out differs: {rel_err(out, o_r):.2e}"
        # the probabilities are a softmax: rows sum to one, and the codes are their correctly rounded quantization
        assert torch.allclose(P[..., :N].sum(-1), torch.ones(B * H, N, device=dev), atol=1e-5)
    # no optional outputs (eval): same codes and output
    out2, qp2, P2, qp16_2, _ = ops.qkr_attn_fwd(qx, qk, qvT, B, N, H, C, se_x, se_k, ctS, scale, se_p, qhi, se_v, v_aft)
    assert P2 is None and qp16_2 is None and torch.equal(out2, out) and torch.equal(qp2, qp)


@pytest.mark.parametrize("C,H,N,B", [(384, 6, 198, 4), (192, 3, 198, 3), (128, 2, 40, 3)])
def test_module_forward_backward_identical_with_and_without_fusion(mods, C, H, N, B):
    ops, Fn = mods
    import ofq_b200.quantization as Q
    from ofq_b200.host.deit import Attention
    torch.manual_seed(5)
    mod = Q.QAttention_qkreparam(Attention(C, H, qkv_bias=True), weight_bits=2, input_bits=2, pretrained_initialized=True)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith(".bias") and p.dim() == 1:
                p.copy_(torch.randn(p.shape) * 0.02)
    mod = mod.cuda().train()
    x0 = torch.randn(B, N, C, device="cuda")
    go = torch.randn(B, N, C, device="cuda")
    with torch.no_grad():
        mod(x0)
    res = {}
    for mode in ("fused", "fused_fwd", "unfused"):
        Fn.FUSED_ATTN, Fn.FUSED_ATTN_BWD = mode != "unfused", mode == "fused"
        try:
            mod.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_(True)
            l0 = ops.LAUNCHES
            y, _ = mod(x)
            nfwd = ops.LAUNCHES - l0
            y.backward(go)
            nall = ops.LAUNCHES - l0
        finally:
            Fn.FUSED_ATTN = Fn.FUSED_ATTN_BWD = True
        res[mode] = (y.detach(), x.grad.clone(), {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None}, nfwd, nall)
    assert res["fused"][3] < res["unfused"][3] and res["fused"][4] < res["fused_fwd"][4]     # the fused kernels really ran
    assert torch.equal(res["fused"][0], res["unfused"][0]) and torch.equal(res["fused_fwd"][0], res["unfused"][0])
    # fused forward + three-kernel backward on the saved probabilities: the same numbers as the unfused path
    assert torch.equal(res["fused_fwd"][1], res["unfused"][1])
    for n in res["unfused"][2]:
        a, b = res["fused_fwd"][2][n], res["unfused"][2][n]
        assert torch.equal(a, b) or rel_err(a, b) < 1e-5, n   # (split-K atomics of the dW GEMMs are order dependent)
    # fused backward: same mathematics, another power-of-two range scale of the 16-bit dS operand (a-priori bound instead of the
    # measured maximum of dP): agreement at the rounding level of the fp16 operands
    gmax = max(v.abs().max().item() for v in res["unfused"][2].values())
    assert rel_err(res["fused"][1], res["unfused"][1]) < 1e-3
    for n in res["unfused"][2]:
        a, b = res["fused"][2][n], res["unfused"][2][n]
        assert rel_err(a, b) < 1e-3 or (a - b).abs().max().item() <= 1e-5 * gmax, f"{n}: {rel_err(a, b):.2e}"
    mod.eval()
    outs = []
    for fused in (True, False):
        Fn.FUSED_ATTN = fused
        try:
            with torch.no_grad():
                outs.append(mod(x0)[0])
        finally:
            Fn.FUSED_ATTN = True
    assert rel_err(outs[0], outs[1]) < 1e-6                   # eval: the unfused path takes the scalar softmax kernel (ulp-level ties)


@pytest.mark.parametrize("C,H,N,B", [(384, 6, 198, 4), (128, 2, 40, 3)])
def test_module_with_quantizing_qkx_epilogue(mods, C, H, N, B):
    """QAttention_qkreparam with the qkx quantizer as the epilogue of the qkx GEMM (ofq_gemm_lsq, fp16 residual plane for the
    backward) against GEMM -> fp32 qkx -> quantizer pass: same codes, so the same output up to the summation order of the logits'
    column term; gradients agree at the rounding level of the fp16 residual (step sizes) and exactly in the mask."""
    ops, Fn = mods
    import ofq_b200.quantization as Q
    from ofq_b200.host.deit import Attention
    torch.manual_seed(6)
    mod = Q.QAttention_qkreparam(Attention(C, H, qkv_bias=True), weight_bits=2, input_bits=2, pretrained_initialized=True)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith(".bias") and p.dim() == 1:
                p.copy_(torch.randn(p.shape) * 0.02)
    mod = mod.cuda().train()
    x0 = torch.randn(B, N, C, device="cuda")
    go = torch.randn(B, N, C, device="cuda")
    with torch.no_grad():
        mod(x0)
    res = {}
    for fused in (True, False):
        Fn.FUSED_QKX = fused
        try:
            mod.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_(True)
            ops.PROFILE = []
            y, _ = mod(x)
            fams = [r[0] for r in ops.PROFILE]
            ops.PROFILE = None
            y.backward(go)
        finally:
            Fn.FUSED_QKX = True
            ops.PROFILE = None
        res[fused] = (y.detach(), x.grad.clone(), {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None}, fams)
    assert "gemm_lsq" in res[True][3] and "gemm_lsq" not in res[False][3]   # the quantizing epilogue really ran
    assert res[True][3].count("lsq_quant") == res[False][3].count("lsq_quant") - 1
    assert rel_err(res[True][0], res[False][0]) < 1e-5
    gmax = max(v.abs().max().item() for v in res[False][2].values())
    assert rel_err(res[True][1], res[False][1]) < 1e-4
    for n in res[False][2]:
        a, b = res[True][2][n], res[False][2][n]
        assert rel_err(a, b) < 1e-4 or (a - b).abs().max().item() <= 1e-5 * gmax, f"{n}: {rel_err(a, b):.2e}"


@pytest.mark.parametrize("B,N,H,bits", [(3, 198, 6, 2), (4, 198, 3, 4), (2, 40, 2, 3), (150, 198, 6, 2), (2, 129, 1, 2)])
def test_fused_backward_against_three_kernel_path(mods, B, N, H, bits):
    """ofq_qkr_attn_bwd (logits recomputed, dP in TMEM, softmax / quantizer backward in the same kernel) against
    dP GEMM -> ofq_softmax_quant_bwd on stored probabilities: dS (un-scaled), its column sums and the step-size gradient."""
    ops, Fn = mods
    from ofq_b200.ops import GEMM_F16, FMT_F16, vec
    C = 64 * H
    torch.manual_seed(B * 77 + N + H)
    dev = "cuda"
    lo, hi = -(2 ** (bits - 1)), 2 ** (bits - 1) - 1
    qhi = 2 ** bits - 1
    qx = torch.randint(lo, hi + 1, (B * N, C), dtype=torch.int8, device=dev)
    qk = torch.randint(lo, hi + 1, (B * N, H * C), dtype=torch.int8, device=dev)
    qv = torch.randint(lo, hi + 1, (B * N, C), dtype=torch.int8, device=dev)
    se_x = torch.rand(N, device=dev) * 0.5 + 0.5
    se_k = (torch.rand(N * H, device=dev) * 0.5 + 0.5) * (2.0 / (C ** 0.5) / max(1, 2 ** (bits - 2)) ** 2)
    ctS = torch.randn(B * N, H, device=dev)
    se_p = torch.rand(N, device=dev) * (0.5 / qhi) + 0.2 / qhi
    sp2 = torch.stack((se_p, 1.0 / se_p)).contiguous()
    se_v = torch.rand(C, device=dev) * 0.1 + 0.05
    sv2 = torch.stack((se_v, 1.0 / se_v)).contiguous()
    v_aft = torch.randn(C, device=dev) * 0.02
    scale, g_p = 64 ** -0.5, 1.0 / ((qhi * B * H * N) ** 0.5)
    dO = torch.randn(B, N, C, device=dev) * 1e-3
    qvT = ops.codes_transpose(qv, B, N, C, C, N * C)
    fwd = ops.qkr_attn_fwd(qx, qk, qvT, B, N, H, C, se_x, se_k, ctS, scale, se_p, qhi, se_v, v_aft, save_p=True, fmt16=FMT_F16,
                           want_rowstat=True)
    out, qp, P, qp16, _, rowstat = fwd
    ldq, ldS = qp.shape[-1], P.shape[-1]
    # ---- three-kernel reference
    amax = torch.zeros(1, device=dev)
    dPq, dvhat = Fn._pv_backward_f16(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, None, qp16, None, amax_dp=amax)
    se_k_hn = se_k.view(N, H).t().contiguous()
    sc = ops.scale_from_max(amax, v1=se_k_hn, v2=se_x, mult=2.0 * scale, product=True)
    dS16_r, _, ldo, colsum_r, ds_r, _ = ops.softmax_quant_bwd(dPq, P, N, H, se_p, qhi, scale, g_p, se_k_hn, True, se_x, fmt=FMT_F16,
                                                               scale4=sc, single=True)
    # ---- fused
    (a16, rowdot, sc_in, qv16), dvhat2 = Fn._pv_backward_f16(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, None, qp16, None, skip_dp=True)
    dS16, ldo2, colsum, ds, sc2 = ops.qkr_attn_bwd(qx, qk, a16, qv16, FMT_F16, B, N, H, C, se_x, se_k, ctS, scale, sp2, qhi, rowstat, rowdot,
                                                   sc_in, se_v, v_aft, -lo, g_p)
    torch.cuda.synchronize()
    assert ldo2 == ldo and torch.equal(dvhat, dvhat2)
    s_r, s_f = float(sc[0]), float(sc2[0])
    assert s_f > 0 and s_f <= s_r * 1.0001 and s_f >= s_r / 4096          # a-priori bound: never tighter than the measured one, a few binades looser
    a = dS16.float()[..., :N] / s_f
    b = dS16_r.float()[..., :N] / s_r
    assert not torch.isinf(dS16.float()).any() and not torch.isnan(a).any()
    assert rel_err(a, b) < 2e-3, rel_err(a, b)                            # two fp16 roundings at different scales
    assert bool((dS16.float()[..., N:] == 0).all())
    assert rel_err(colsum, colsum_r) < 1e-4
    assert rel_err(ds, ds_r) < 1e-4
