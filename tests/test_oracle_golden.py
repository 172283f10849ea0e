"""Pin oracle/ofq_oracle.py against golden vectors produced by the unmodified reference on CPU
(tests/golden/make_golden.py). Forward values must be bit-identical (same op sequence in fp32);
gradients are compared at 1e-6 because autograd may reorder accumulations."""
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import ofq_oracle as O


def params_from(d, requires_grad=True):
    P = {}
    for k, v in d.items():
        if k.startswith("param."):
            t = v.clone()
            if requires_grad and t.is_floating_point():
                t.requires_grad_(True)
            P[k[len("param."):]] = t
    return P


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("name", ["kat", "rand", "dyadic"])
def test_statsq(bits, name):
    g = load_golden("statsq")
    w = g[f"{name}.w"].clone().requires_grad_(True)
    y = O.statsq(w, bits)
    assert torch.equal(y, g[f"{name}.b{bits}.out"])
    assert torch.equal(y, g[f"{name}.b{bits}.out_cga"])  # the _4_qkreparam_cga quantizer is value-identical
    y.backward(torch.ones_like(y) * 0.5)
    assert torch.equal(w.grad, g[f"{name}.b{bits}.grad"])
    assert torch.equal(O.statsq_scale(w).squeeze(), g[f"{name}.b{bits}.s"].squeeze())
    # integer view reproduces the fake-quant floats exactly
    codes, sf = O.statsq_codes(w.detach(), bits)
    n = float(2 ** (bits - 1))
    # (the reference returns (wq - w) + w, which is wq only up to one rounding of w: statsq.py:148)
    assert torch.allclose(sf * (codes.float() / (2 * n)), y.detach(), rtol=1e-6, atol=1e-9)
    assert int(codes.abs().max()) <= 2 ** bits - 1 and bool((codes % 2 != 0).all())


def test_statsq_kat_codes():
    """SURVEY.md §8a row 4 known-answer row."""
    w = load_golden("statsq")["kat.w"]
    assert O.statsq_scale(w).item() == pytest.approx(0.7625)
    assert O.statsq_codes(w, 2)[0].tolist() == [[1, -1, 1, -3, 1, 1, 3, -3]]
    assert O.statsq_codes(w, 3)[0].tolist() == [[1, -3, 3, -5, 1, 1, 7, -7]]
    assert O.statsq_codes(w, 4)[0].tolist() == [[3, -5, 7, -9, 1, 1, 15, -15]]


@pytest.mark.parametrize("bit", [2, 3, 4])
@pytest.mark.parametrize("pos", [False, True])
def test_lsq_rows_cols(bit, pos):
    g = load_golden("lsq")
    tag = f"b{bit}.{'u' if pos else 's'}"
    _, hi = O.lsq_levels(bit, pos)
    x = g["x3"].abs() if pos else g["x3"]
    assert torch.equal(O.lsq_init_rows(x, hi, pos), g[f"rows3.{tag}.s_init"])
    assert torch.equal(O.lsq_init_cols(x, hi, pos), g[f"cols3.{tag}.s_init"])
    for kind, fn in (("rows3", O.lsq_rows), ("cols3", O.lsq_cols)):
        xin = x.clone().requires_grad_(True)
        s = g[f"{kind}.{tag}.s"].clone().requires_grad_(True)
        y = fn(xin, s, bit, pos)
        assert torch.equal(y, g[f"{kind}.{tag}.out"])
        y.backward(g[f"rows3.{tag}.go"])
        assert torch.equal(xin.grad, g[f"{kind}.{tag}.dx"])
        assert rel_err(s.grad, g[f"{kind}.{tag}.ds"]) < 1e-6


@pytest.mark.parametrize("bit", [2, 3, 4])
def test_lsq_probabilities(bit):
    g = load_golden("lsq")
    xin = g["x4"].clone().requires_grad_(True)
    s = g[f"rows4.b{bit}.s"].clone().requires_grad_(True)
    y = O.lsq_rows(xin, s, bit, True)
    assert torch.equal(y, g[f"rows4.b{bit}.out"])
    y.backward(g[f"rows4.b{bit}.go"])
    assert torch.equal(xin.grad, g[f"rows4.b{bit}.dx"])
    assert rel_err(s.grad, g[f"rows4.b{bit}.ds"]) < 1e-6


def _check_layer(name, fn):
    g = load_golden(name)
    P = params_from(g)
    x = g["x"].clone().requires_grad_(True)
    y = fn(x, P)
    assert torch.equal(y, g["out"]), f"{name}: forward differs, rel {rel_err(y, g['out']):.3e}"
    y.backward(g["go"])
    assert rel_err(x.grad, g["dx"]) < 1e-6
    checked = 0
    for k, v in g.items():
        if k.startswith("grad."):
            p = P[k[len("grad."):]]
            assert p.grad is not None, k
            scale = max(v.abs().max().item(), 1e-12)
            assert (p.grad - v).abs().max().item() <= 2e-5 * scale + 1e-9, k
            checked += 1
    assert checked > 0


@pytest.mark.parametrize("bits", [2, 4])
def test_qlinear(bits):
    _check_layer(f"qlinear_w{bits}a{bits}", lambda x, P: O.qlinear(x, P, "", bits, bits))


@pytest.mark.parametrize("bits", [2, 4])
def test_qmlp(bits):
    _check_layer(f"qmlp_w{bits}a{bits}", lambda x, P: O.qmlp(x, P, "", bits, bits))


@pytest.mark.parametrize("bits", [2, 4])
def test_qattention(bits):
    _check_layer(f"qattention_w{bits}a{bits}", lambda x, P: O.qattention(x, P, "", 2, bits, bits))


@pytest.mark.parametrize("bits", [2, 4])
def test_qattention_qkr(bits):
    _check_layer(f"qattention_qkr_w{bits}a{bits}", lambda x, P: O.qattention_qkr(x, P, "", 2, bits, bits))
    assert load_golden(f"qattention_qkr_w{bits}a{bits}")["cga_variant_identical"].item() == 1.0


@pytest.mark.parametrize("qkr", [False, True])
def test_deit_step(qkr):
    g = load_golden(f"deit_tiny2_{'qkr' if qkr else 'plain'}_w2a2")
    P = params_from(g)
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["img_seed"])))
    state = {"signed": int(g["param.patch_embed.proj.input_quant_fn.signed"].item())}
    cls, dist = O.deit_forward(img, P, depth=2, heads=2, wbits=2, abits=2, qkr=qkr, state=state)
    assert torch.equal(cls, g["cls"]) and torch.equal(dist, g["dist"])
    loss = F.cross_entropy(cls, g["labels"]) + F.cross_entropy(dist, g["labels"])
    assert torch.equal(loss, g["loss"])
    loss.backward()
    n = 0
    for k, v in g.items():
        if not k.startswith("gnorm."):
            continue
        name = k[len("gnorm."):]
        gr = P[name].grad
        assert gr is not None, name
        assert abs(gr.norm().item() - v.item()) <= 1e-4 * v.item() + 1e-10, name
        ref = g["grad." + name]
        mine = gr if gr.numel() <= 4096 else gr.flatten()[:: max(1, gr.numel() // 2048)][:2048]
        scale = max(ref.abs().max().item(), 1e-12)
        assert (mine - ref).abs().max().item() <= 1e-4 * scale + 1e-10, name
        n += 1
    assert n > 50
    with torch.no_grad():
        ev = O.deit_forward(img, P, depth=2, heads=2, wbits=2, abits=2, qkr=qkr, state=state, training=False)
    assert torch.equal(ev, g["eval_logits"])


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("br", [0.005, 0.05])
def test_cga_mask(bits, br):
    g = load_golden("cga")
    m = O.cga_freeze_mask(g["w"], bits, br)
    assert torch.equal(m, g[f"mask.b{bits}.br{br}"])
    assert 0 < (m == 0).sum().item() < m.numel()


def test_cga_masked_adamw():
    g = load_golden("cga")
    w = g["w"].clone()
    m = torch.zeros_like(w)
    v = torch.zeros_like(w)
    for step in range(3):
        f = O.cga_masked_step(w, g[f"step{step}.grad"], m, v, step + 1, lr=1e-3, wd=0.05, bits=2, boundary_range=0.05)
        assert torch.equal(f, g[f"step{step}.mask"])
        frozen = f == 1
        assert torch.equal(w[frozen], g[f"step{step}.w"][frozen])          # frozen weights: bit-identical
        assert rel_err(w, g[f"step{step}.w"]) < 1e-7
        assert rel_err(m, g[f"step{step}.exp_avg"]) < 1e-6
        assert rel_err(v, g[f"step{step}.exp_avg_sq"]) < 1e-6


@pytest.mark.parametrize("shift", [0, 3])
@pytest.mark.parametrize("qkr", [False, True])
def test_swin_window_attention(shift, qkr):
    name = f"qattention_swin_{'qkr' if qkr else 'plain'}_shift{shift}_w3a3"
    _check_layer(name, lambda x, P: O.swin_window_attention(x, P, "", 2, 3, 3, qkr, (7, 7), (shift, shift)))


@pytest.mark.parametrize("qkr", [False, True])
def test_swin_step(qkr):
    g = load_golden(f"swin_tiny2_{'qkr' if qkr else 'plain'}_w3a3")
    P = params_from(g)
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["img_seed"])))
    state = {"signed": int(g["param.features.0.0.input_quant_fn.signed"].item())}
    logits = O.swin_forward(img, P, (2, 2), (1, 2), 3, 3, qkr, state)
    assert torch.equal(logits, g["logits"])
    loss = F.cross_entropy(logits, g["labels"])
    assert torch.equal(loss, g["loss"])
    loss.backward()
    n = 0
    for k, v in g.items():
        if not k.startswith("gnorm."):
            continue
        name = k[len("gnorm."):]
        gr = P[name].grad
        assert gr is not None, name
        ref = g["grad." + name]
        mine = gr if gr.numel() <= 4096 else gr.flatten()[:: max(1, gr.numel() // 2048)][:2048]
        scale = max(ref.abs().max().item(), 1e-12)
        assert (mine - ref).abs().max().item() <= 1e-4 * scale + 1e-10, name
        n += 1
    assert n > 80


# ------------------------------------------------------------------------------------------------ KD losses (f2)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_kd_losses_match_reference_classes(tag):
    """oracle.kl_loss_soft / kd_loss_soft_and_hard against the reference's KLLossSoft / KDLossSoftandHard
    (src/quantization/utils.py:44-77), values and gradients (tests/golden/make_golden_kd.py)."""
    g = load_golden("kd_loss")
    cls = g[f"{tag}.cls"].clone().requires_grad_(True)
    dist = g[f"{tag}.dist"].clone().requires_grad_(True)
    teacher, teacher_dist, y = g[f"{tag}.teacher"], g[f"{tag}.teacher_dist"], g[f"{tag}.y"].long()
    loss = O.kd_loss_soft_and_hard((cls, dist), y, (teacher, teacher_dist))
    loss.backward()
    assert torch.equal(loss.detach(), g[f"{tag}.sh_tuple.loss"])
    assert torch.equal(cls.grad, g[f"{tag}.sh_tuple.dcls"]) and torch.equal(dist.grad, g[f"{tag}.sh_tuple.ddist"])
    cls.grad = None
    loss = O.kd_loss_soft_and_hard(cls, y, teacher)
    loss.backward()
    assert torch.equal(loss.detach(), g[f"{tag}.sh_single.loss"]) and torch.equal(cls.grad, g[f"{tag}.sh_single.dcls"])
    for T in (1.0, 2.5):
        cls.grad = None
        loss = O.kl_loss_soft((cls, dist), (teacher, teacher_dist), T=T)
        loss.backward()
        assert torch.equal(loss.detach(), g[f"{tag}.soft_T{T}.loss"]) and torch.equal(cls.grad, g[f"{tag}.soft_T{T}.dcls"])
