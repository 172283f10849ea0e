"""Parity at BASELINE.json's full sizes (DeiT-S, batch 128: M = 25344 rows) through size-independent properties and
same-op torch-on-GPU cross-checks; the CPU oracle is too slow at these sizes."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
M, C, H, N = 25344, 384, 6, 198


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from ofq_b200 import _lib, ops
    assert _lib.load().ofq_device_ok() == 1
    return ops


def test_statsq_codes_monotone_and_in_range_fullsize(ops):
    """Within every output channel the code is a non-decreasing function of the weight (sortedness), odd, in range,
    and the scale is 2*mean|w| to fp32 round-off — on every quantized weight shape of a DeiT-S QKR block."""
    torch.manual_seed(0)
    for rows, cols in ((384, 384), (1536, 384), (384, 1536), (2304, 384)):
        w = torch.nn.init.trunc_normal_(torch.empty(rows, cols, device="cuda"), std=0.02)
        for bits in (2, 3, 4):
            codes, colscale, sf, _, _ = ops.statsq_codes(w, bits)
            c = codes.int()
            assert bool((c % 2 != 0).all()) and int(c.abs().max()) == 2 ** bits - 1
            order = torch.argsort(w, dim=1)
            assert bool((torch.gather(c, 1, order).diff(dim=1) >= 0).all())
            assert rel_err(sf, 2 * w.abs().mean(1)) < 1e-6
            assert torch.equal(colscale, sf / 2 ** bits)


def test_lsq_codes_equal_torch_same_ops_fullsize(ops):
    """25344 x 1536 activations (fc2 input): codes == round(clamp((x+b)/s)) evaluated by torch on the same GPU."""
    torch.manual_seed(1)
    x = F.gelu(torch.randn(M, 4 * C, device="cuda"))
    b4 = torch.randn(4 * C, device="cuda") * 0.02
    s = torch.rand(N, device="cuda") * 0.2 + 0.05
    se = ops.lsq_effective_scale(s, 1.0 / ((3 * 128 * 4 * C) ** 0.5))
    codes = ops.lsq_quant(x, b4, se, ops.PER_ROW, N, 1, 0, 3)
    ref = torch.round(torch.clamp((x + b4) / se.repeat(M // N).unsqueeze(1), 0, 3)).to(torch.int8)
    assert torch.equal(codes, ref)
    assert int(codes.max()) == 3 and int(codes.min()) == 0


def test_qlinear_gemm_exact_and_linear_fullsize(ops):
    """fc1-shaped int8 GEMM (25344 x 1536 x 384): exact integer accumulation and linearity in the weight codes."""
    torch.manual_seed(2)
    A = torch.randint(-2, 2, (M, C), dtype=torch.int8, device="cuda")
    B1 = torch.randint(-3, 4, (4 * C, C), dtype=torch.int8, device="cuda")
    B2 = torch.randint(-3, 4, (4 * C, C), dtype=torch.int8, device="cuda")
    outs = []
    for Bm in (B1, B2, B1 + B2):
        out = torch.empty(M, 4 * C, device="cuda")
        ops.gemm(ops.GEMM_I8, A, (C, 0, 0, 0), Bm, (C, 0, 0, 0), out, (4 * C, 0, 0), M, 4 * C, C)
        outs.append(out)
    assert torch.equal(outs[0] + outs[1], outs[2])                        # integers < 2^24: fp32 sums are exact
    ref = (A[:4096].float() @ B1.float().T)                               # fp32 matmul of small ints is exact
    assert torch.equal(outs[0][:4096], ref)


def test_attention_probability_rows_fullsize(ops):
    """128 x 6 heads x 198 x 198 scores: every probability row sums to 1, codes never exceed the level count and
    rowsum == s * sum(codes)."""
    torch.manual_seed(3)
    S = torch.randn(128 * H, N, 200, device="cuda")
    se = torch.full((N,), 0.01, device="cuda")
    P, codes, rowsum = ops.softmax_quant(S, N, H, se, 3)
    assert torch.allclose(P[..., :N].sum(-1), torch.ones(128 * H, N, device="cuda"), atol=1e-5)
    assert int(codes.max()) <= 3 and int(codes.min()) >= 0 and bool((codes[..., N:] == 0).all())
    assert torch.allclose(rowsum, codes[..., :N].float().sum(-1) * 0.01, rtol=1e-6)


def test_cga_step_fullsize_properties(ops):
    """Masked AdamW on every masked weight shape of DeiT-S: mask == the cga.py formula restated with torch ops on the
    GPU, ~99 % frozen at BR = 0.005 (SURVEY.md §7 hard part 8), frozen weights bit-identical, moments as AdamW(g=0)."""
    torch.manual_seed(4)
    for rows, cols in ((384, 384), (1536, 384), (384, 1536)):
        w = torch.nn.init.trunc_normal_(torch.empty(rows, cols, device="cuda"), std=0.02)
        g = torch.randn_like(w) * 1e-3
        m = torch.rand_like(w) * 1e-3
        v = torch.rand_like(w) * 1e-6
        w0, m0, v0 = w.clone(), m.clone(), v.clone()
        mask = torch.empty(w.shape, dtype=torch.uint8, device="cuda")
        ops.cga_adamw_(w, g, m, v, 7, 1e-5, 0.9, 0.999, 1e-8, 0.05, bits=2, boundary_range=0.005, mask_out=mask)
        sf = 2 * w0.abs().mean(1, keepdim=True)
        b4 = torch.clamp(w0 / sf, -1.0, 1.0 - 1e-6) * 2.0 - 0.5
        r = torch.round(b4)
        nf = torch.zeros_like(w0)
        for i in range(int(r.min()), int(r.max())):
            d = b4 - float(i)
            nf += ((d <= 0.505) & (d >= 0.495)).float()
        ref_mask = (1 - nf).to(torch.uint8)
        agree = (mask == ref_mask).float().mean().item()
        assert agree > 1 - 1e-5                                            # only scale-ulp ties may differ
        frozen = mask.bool()
        assert 0.98 < frozen.float().mean().item() < 0.999
        assert torch.equal(w[frozen], w0[frozen])
        assert torch.allclose(m[frozen], m0[frozen] * 0.9, rtol=1e-6) and torch.allclose(v[frozen], v0[frozen] * 0.999, rtol=1e-6)
        assert not torch.equal(w[~frozen], w0[~frozen])


def test_deit_small_step_deterministic_forward(ops):
    """The full DeiT-S W2A2 QKR model at batch 32: the forward has no atomics, so two runs give a bit-identical loss;
    gradients (split-K / reduce-add order) agree to fp32 round-off."""
    import ofq_b200.quantization as Q
    from ofq_b200.host.deit import deit_small_distilled_patch16_224
    torch.manual_seed(5)
    model = deit_small_distilled_patch16_224(num_classes=1000)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(Q.deit_qmodule_names(12), 2, 2), pretrained_initialized=True,
                                             qk_reparam=True).cuda()
    img = torch.randn(32, 3, 224, 224, device="cuda")
    lbl = torch.randint(0, 1000, (32,), device="cuda")
    model.eval()
    with torch.no_grad():
        model(img)
    model.train()
    losses, grads = [], []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        (c, d), _ = model(img)
        loss = F.cross_entropy(c, lbl) + F.cross_entropy(d, lbl)
        loss.backward()
        losses.append(loss.detach().clone())
        grads.append(torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]))
    assert torch.equal(losses[0], losses[1]) and torch.isfinite(losses[0])
    assert rel_err(grads[0], grads[1]) < 1e-5
