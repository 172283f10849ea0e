"""GPU parity of the drop-in modules against golden outputs of the UNMODIFIED reference (tests/golden/*.npz,
generated on CPU in fp32): forward outputs / logits / loss and every parameter gradient.

Tolerance: BASELINE.json asks for <= 1e-3 relative on outputs, logits, loss and scale gradients; we assert a
tighter 1e-4 on outputs and 1e-3 on gradients (norm-wise relative error ||a-b||/||b||). Gradients that are
analytically zero in the reference (shift added to an operand whose contribution cancels in softmax:
move_k_aft / move_qkx_aft) are checked against an absolute bound instead."""
import os
from functools import partial

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
OUT_TOL, GRAD_TOL = 1e-4, 1e-3
ANALYTIC_ZERO = ("move_k_aft.bias", "move_qkx_aft.bias")


@pytest.fixture(scope="module")
def Q():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import ofq_b200.quantization as Q
    from ofq_b200 import _lib
    assert _lib.load().ofq_device_ok() == 1
    return Q


def load_params(mod, g):
    sd = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    missing, unexpected = mod.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    missing = [m for m in missing if not m.endswith("relative_position_index")]   # a constant buffer, not stored in goldens
    assert not missing, missing          # lazily created LSQ scales must load from a checkpoint too
    return mod


REPORT = os.environ.get("OFQ_GRAD_REPORT")      # diagnostic: append "<rel err> <abs/gmax> <name>" lines instead of asserting


def check_grads(named_params, g, sampled=False):
    params = dict(named_params)
    gmax = max(v.abs().max().item() for k, v in g.items() if k.startswith("grad."))
    n = 0
    if REPORT:
        with open(REPORT, "a") as fh:
            fh.write(f"# {os.environ.get('PYTEST_CURRENT_TEST', '')}\n")
            for k, ref in g.items():
                if k.startswith("grad."):
                    mine = params[k[5:]].grad.detach().cpu()
                    fh.write(f"{rel_err(mine, ref):.2e} {(mine - ref).abs().max().item() / gmax:.2e} {k[5:]}\n")
        return 1
    for k, ref in g.items():
        if not k.startswith("grad."):
            continue
        name = k[len("grad."):]
        p = params[name]
        assert p.grad is not None, name
        mine = p.grad.detach().cpu()
        if sampled and mine.numel() > 4096:
            mine = mine.flatten()[:: max(1, mine.numel() // 2048)][:2048]
        if name.endswith(ANALYTIC_ZERO):
            assert mine.abs().max().item() <= 1e-5 * gmax, name
        else:
            # gradients that are pure round-off in the reference (|g| < 1e-5 of the largest gradient, e.g. move_*_b4
            # of an operand that never clips at 4 bit) are held to the same absolute bound
            ok = rel_err(mine, ref) < GRAD_TOL or (mine - ref).abs().max().item() <= 1e-5 * gmax
            assert ok, f"{name}: {rel_err(mine, ref):.2e}"
        n += 1
    return n


def run_layer(mod, g):
    mod = load_params(mod, g).cuda().train()
    x = g["x"].cuda().requires_grad_(True)
    y = mod(x)
    y = y[0] if isinstance(y, tuple) else y
    assert rel_err(y.detach().cpu(), g["out"]) < OUT_TOL
    y.backward(g["go"].cuda())
    assert rel_err(x.grad.cpu(), g["dx"]) < GRAD_TOL
    assert check_grads(mod.named_parameters(), g) > 0


@pytest.mark.parametrize("bits", [2, 4])
def test_qlinear(Q, bits):
    run_layer(Q.QLinear(m=nn.Linear(32, 48), weight_bits=bits, input_bits=bits), load_golden(f"qlinear_w{bits}a{bits}"))


@pytest.mark.parametrize("bits", [2, 4])
def test_qmlp(Q, bits):
    from ofq_b200.host.deit import Mlp
    run_layer(Q.QMLP(m=Mlp(32, 128), weight_bits=bits, input_bits=bits), load_golden(f"qmlp_w{bits}a{bits}"))


@pytest.mark.parametrize("bits", [2, 4])
def test_qattention(Q, bits):
    from ofq_b200.host.deit import Attention
    run_layer(Q.QAttention(Attention(32, 2, qkv_bias=True), weight_bits=bits, input_bits=bits),
              load_golden(f"qattention_w{bits}a{bits}"))


@pytest.mark.parametrize("bits", [2, 4])
@pytest.mark.parametrize("cga", [False, True])
def test_qattention_qkreparam(Q, bits, cga):
    from ofq_b200.host.deit import Attention
    cls = Q.QAttention_qkreparam_4_cga if cga else Q.QAttention_qkreparam
    kw = {"boundaryRange": 0.005} if cga else {}
    run_layer(cls(Attention(32, 2, qkv_bias=True), weight_bits=bits, input_bits=bits, **kw),
              load_golden(f"qattention_qkr_w{bits}a{bits}"))


def test_qattention_qkreparam_fused_fp16_operands(Q):
    """C = 128 (a multiple of the 128-column streaming group): the backward takes the fused route of the fp16 mode: GEMM
    epilogues track max |d v_hat| / |d k_hat|, the LSQ backward passes write the fp16 GEMM operands directly (no fp32
    d v_out / d qkx, no ofq_grad_prep). Checked against the CPU oracle (autograd of the reference op sequence) on the same
    parameters, and against the unfused route."""
    from oracle import ofq_oracle as O
    from ofq_b200.host.deit import Attention
    from ofq_b200.quantization import functional as Fn
    torch.manual_seed(31)
    B, N, C, H, bits = 3, 40, 128, 2, 2
    mod = Q.QAttention_qkreparam(Attention(C, H, qkv_bias=True), weight_bits=bits, input_bits=bits, pretrained_initialized=True)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith(".bias") and p.dim() == 1:
                p.copy_(torch.randn(p.shape) * 0.02)
    mod = mod.cuda().train()
    x0 = torch.randn(B, N, C)
    go = torch.randn(B, N, C)
    with torch.no_grad():
        mod(x0.cuda())                                   # creates the LSQ step sizes
    grads = {}
    for fused in (True, False):
        Fn.FUSED16 = fused
        try:
            mod.zero_grad()
            x = x0.cuda().requires_grad_(True)
            y, _ = mod(x)
            y.backward(go.cuda())
        finally:
            Fn.FUSED16 = True
        grads[fused] = {n: p.grad.detach().cpu().clone() for n, p in mod.named_parameters() if p.grad is not None}
        grads[fused]["x"] = x.grad.detach().cpu().clone()
        grads[fused]["y"] = y.detach().cpu()
    # oracle on the same parameters
    P = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}
    xo = x0.clone().requires_grad_(True)
    yo = O.qattention_qkr(xo, P, "", H, bits, bits)
    yo.backward(go)
    assert rel_err(grads[True]["y"], yo.detach()) < OUT_TOL
    gmax = max(P[n].grad.abs().max().item() for n in grads[True] if n in P and P[n].grad is not None)
    for n, g in grads[True].items():
        if n == "y":
            continue
        ref = xo.grad if n == "x" else P[n].grad
        if n.endswith(ANALYTIC_ZERO):
            assert g.abs().max().item() <= 1e-5 * gmax, n
            continue
        assert rel_err(g, ref) < GRAD_TOL or (g - ref).abs().max().item() <= 1e-5 * gmax, f"{n}: {rel_err(g, ref):.2e}"
        assert rel_err(g, grads[False][n]) < GRAD_TOL or (g - grads[False][n]).abs().max().item() <= 1e-5 * gmax, n


def test_unknown_quant_method_raises(Q):
    with pytest.raises(ValueError, match="Unknown quant_method"):
        Q.QLinear(m=nn.Linear(8, 8), weight_quant_method="lsq")


def test_lazy_scale_init_matches_reference_formula(Q):
    """First forward creates `s` from that batch (lsq.py:544-569): 2*mean|x + b4| / sqrt(thd_pos) per token."""
    torch.manual_seed(0)
    lin = Q.QLinear(m=nn.Linear(64, 32), weight_bits=2, input_bits=2).cuda()
    assert lin.input_quant_fn.s is None and "input_quant_fn.s" not in lin.state_dict()
    x = torch.randn(4, 10, 64, device="cuda")
    lin.eval()
    with torch.no_grad():
        lin(x)
    ref = 2 * (x + lin.move_b4.bias).abs().mean(-1).mean(0) / (1 ** 0.5)
    assert torch.allclose(lin.input_quant_fn.s, ref, rtol=1e-6)
    assert "input_quant_fn.s" in lin.state_dict() and lin.input_quant_fn.s.requires_grad


@pytest.mark.parametrize("qkr", [False, True])
def test_deit_step_against_reference(Q, qkr):
    """Depth-2, width-64 distilled DeiT, every qmodule of configs/ours_imagenet_recipe.attn_q.yml quantized, W2A2."""
    from ofq_b200.host.deit import DistilledVisionTransformer
    g = load_golden(f"deit_tiny2_{'qkr' if qkr else 'plain'}_w2a2")
    model = DistilledVisionTransformer(embed_dim=64, depth=2, num_heads=2, num_classes=10)
    names = Q.deit_qmodule_names(2)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(names, 2, 2), pretrained_initialized=True, qk_reparam=qkr)
    model = load_params(model, g).cuda().train()
    ref_keys = {k[len("param."):] for k in g if k.startswith("param.")}
    assert set(model.state_dict().keys()) == ref_keys            # checkpoint-compatible key set
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["img_seed"]))).cuda()
    labels = g["labels"].cuda()
    (cls, dist), _ = model(img)
    assert rel_err(cls.detach().cpu(), g["cls"]) < OUT_TOL and rel_err(dist.detach().cpu(), g["dist"]) < OUT_TOL
    loss = F.cross_entropy(cls, labels) + F.cross_entropy(dist, labels)
    assert abs(loss.item() - g["loss"].item()) <= OUT_TOL * abs(g["loss"].item())
    loss.backward()
    assert check_grads(model.named_parameters(), g, sampled=True) > 50
    model.eval()
    with torch.no_grad():
        ev, _ = model(img)
    assert rel_err(ev.cpu(), g["eval_logits"]) < OUT_TOL


@pytest.mark.parametrize("shift", [0, 3])
@pytest.mark.parametrize("qkr", [False, True])
def test_swin_window_attention(Q, shift, qkr):
    """Quantized shifted-window attention (relative-position bias + 0/-100 shift mask), W3A3, vs the reference."""
    from ofq_b200.host.swin import ShiftedWindowAttention
    cls = Q.QAttention_swin_qkreparam if qkr else Q.QAttention_swin
    mod = cls(ShiftedWindowAttention(32, [7, 7], [shift, shift], 2), weight_bits=3, input_bits=3)
    run_layer(mod, load_golden(f"qattention_swin_{'qkr' if qkr else 'plain'}_shift{shift}_w3a3"))


@pytest.mark.parametrize("shift", [0, 3])
def test_swin_window_attention_4_cga_class(Q, shift):
    """QAttention_swin_qkreparam_4_cga (swin_attention_and_mlp.py:554-671): the CGA weight quantizer is value- and
    gradient-identical to StatsQ, so the class must reproduce the QKR golden of the reference."""
    from ofq_b200.host.swin import ShiftedWindowAttention
    mod = Q.QAttention_swin_qkreparam_4_cga(ShiftedWindowAttention(32, [7, 7], [shift, shift], 2), weight_bits=3, input_bits=3,
                                            boundaryRange=0.005)
    run_layer(mod, load_golden(f"qattention_swin_qkr_shift{shift}_w3a3"))


def test_patch_embed_against_reference_golden(Q):
    """LSQ_QConv2d (8-bit patch embedding, qlinear.py:138-191) on the int8 tensor-core path against the REFERENCE's own module
    (tests/golden/make_golden_ends.py): output, input gradient and every parameter gradient."""
    from ofq_b200.quantization.modules import qlinear as QL
    run_layer(QL.LSQ_QConv2d(m=nn.Conv2d(3, 64, 16, 16), pretrained_initialized=True), load_golden("qconv2d_patch16"))


def test_head_against_reference_golden(Q):
    """LSQ_QLinear4head (8-bit classifier head, qlinear.py:193-238) on the int8 tensor-core path (functional.HeadLinearFn) against
    the REFERENCE's own module."""
    from ofq_b200 import ops
    from ofq_b200.quantization.modules import qlinear as QL
    mod = QL.LSQ_QLinear4head(m=nn.Linear(192, 1000), weight_quant_method="lsq", pretrained_initialized=True)
    l0 = ops.LAUNCHES
    run_layer(mod, load_golden("qlinear4head"))
    assert ops.LAUNCHES > l0                           # the step sizes came from the checkpoint: the native path ran


@pytest.mark.parametrize("qkr", [False, True])
def test_swin_step_against_reference(Q, qkr):
    """Two-stage Swin (depths 2+2, width 32/64, 7x7 windows, patch merging `reduction` QLinear), W3A3."""
    from ofq_b200.host.swin import SwinTransformer
    g = load_golden(f"swin_tiny2_{'qkr' if qkr else 'plain'}_w3a3")
    model = SwinTransformer(embed_dim=32, depths=(2, 2), num_heads=(1, 2), num_classes=10)
    names = Q.swin_qmodule_names((2, 2))
    model = Q.replace_module_by_qmodule_swin(model, Q.make_qconfigs(names, 3, 3), pretrained_initialized=True, qk_reparam=qkr)
    model = load_params(model, g).cuda().train()
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["img_seed"]))).cuda()
    logits, _ = model(img)
    assert rel_err(logits.detach().cpu(), g["logits"]) < OUT_TOL
    loss = F.cross_entropy(logits, g["labels"].cuda())
    assert abs(loss.item() - g["loss"].item()) <= OUT_TOL * abs(g["loss"].item())
    loss.backward()
    assert check_grads(model.named_parameters(), g, sampled=True) > 80


def test_step_prologue_matches_per_layer_path(Q):
    """The step prologue (all StatsQ codes, W_qk products and LSQ step sizes of the model in three launches per forward,
    ofq_b200/prologue.py) is a pure re-scheduling: logits, loss and every gradient are bit-identical to the per-layer path,
    also after the weights changed (an optimizer step) and for a different batch size (new gradient-scale factors)."""
    from ofq_b200 import ops, prologue
    from ofq_b200.cga import CGAAdamW, param_groups_weight_decay
    from ofq_b200.host.deit import DistilledVisionTransformer
    torch.manual_seed(41)
    depth = 2
    model = DistilledVisionTransformer(embed_dim=128, depth=depth, num_heads=2, num_classes=10)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(Q.deit_qmodule_names(depth), 2, 2),
                                             pretrained_initialized=True, qk_reparam=True).cuda()
    img = torch.randn(3, 3, 224, 224, device="cuda")
    lbl = torch.tensor([1, 5, 7], device="cuda")
    model.eval()
    with torch.no_grad():
        model(img)
    model.train()
    opt = CGAAdamW(param_groups_weight_decay(model, 0.05, model.no_weight_decay()), lr=1e-3)

    def run(x, y):
        model.zero_grad(set_to_none=True)
        (cls, dst), _ = model(x)
        loss = F.cross_entropy(cls, y) + F.cross_entropy(dst, y)
        loss.backward()
        return cls.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    assert prologue.ENABLED and getattr(model, "_ofq_prologue", None) is not None
    for x, y in ((img, lbl), (img[:2], lbl[:2])):
        run(x, y)                                   # registers the jobs of this batch shape
        l0 = ops.LAUNCHES
        cls_a, g_a = run(x, y)                      # served by the prologue
        n_pro = ops.LAUNCHES - l0
        assert model._ofq_prologue.fresh is False and len(model._ofq_prologue.statsq) >= 5 * depth
        prologue.ENABLED = False
        try:
            l0 = ops.LAUNCHES
            cls_b, g_b = run(x, y)                  # per-layer path
            n_ref = ops.LAUNCHES - l0
        finally:
            prologue.ENABLED = True
        assert n_pro < n_ref - 10 * depth           # ~13 small launches per block became 3 per model
        assert torch.equal(cls_a, cls_b)
        gmax = max(v.abs().max().item() for v in g_b.values())
        for n in g_b:
            assert rel_err(g_a[n], g_b[n]) < 1e-5 or (g_a[n] - g_b[n]).abs().max().item() <= 1e-6 * gmax, n
        opt.step()                                  # weights move: the next forward must see the new codes


def test_prologue_inference_cache_and_stale_graph_detection(Q):
    """(1) Inference with unchanged weights re-uses the prologue's persistent buffers (no StatsQ / W_qk / step-size launches);
    an optimizer step (CGAAdamW writes through raw pointers and bumps prologue.WEIGHT_EPOCH; torch optimizers bump the
    version counters) invalidates it. (2) A graph whose saved weight codes were overwritten by a later forward after the
    weights changed refuses to run its backward instead of silently using the new codes."""
    from ofq_b200 import ops
    from ofq_b200.cga import CGAAdamW, param_groups_weight_decay
    from ofq_b200.host.deit import DistilledVisionTransformer
    torch.manual_seed(45)
    depth = 2
    model = DistilledVisionTransformer(embed_dim=128, depth=depth, num_heads=2, num_classes=10)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(Q.deit_qmodule_names(depth), 2, 2),
                                             pretrained_initialized=True, qk_reparam=True).cuda()
    img = torch.randn(3, 3, 224, 224, device="cuda")
    lbl = torch.tensor([1, 5, 7], device="cuda")
    model.eval()
    pro = model._ofq_prologue

    def infer():
        l0 = ops.LAUNCHES
        with torch.no_grad():
            out = model(img)[0].clone()
        return out, ops.LAUNCHES - l0

    infer(); infer()                                   # creates the step sizes, registers the jobs
    out_a, n_a = infer()
    hits = pro.cache_hits
    out_b, n_b = infer()
    assert pro.cache_hits == hits + 1 and n_b <= n_a and torch.equal(out_a, out_b)
    # a training step changes the weights: the next inference re-produces the codes and sees the new weights
    model.train()
    opt = CGAAdamW(param_groups_weight_decay(model, 0.05, model.no_weight_decay()), lr=1e-2)
    (cls, dst), _ = model(img)
    (F.cross_entropy(cls, lbl) + F.cross_entropy(dst, lbl)).backward()
    opt.step()
    model.eval()
    hits = pro.cache_hits
    out_c, n_c = infer()
    assert pro.cache_hits == hits and n_c > n_b and not torch.equal(out_c, out_b)
    infer()                                            # (drops the training-mode twins of the jobs: tables rebuilt once more)
    hits = pro.cache_hits
    out_d, _ = infer()
    assert pro.cache_hits == hits + 1 and torch.equal(out_c, out_d)
    with torch.no_grad():                              # a torch-side in-place update is seen through the version counters
        model.blocks[0].mlp.fc1.weight.mul_(1.01)
    hits = pro.cache_hits
    infer()
    assert pro.cache_hits == hits
    # stale graph: forward, optimizer step, ANOTHER forward (re-produces the buffers), then the first graph's backward
    model.train()
    (cls, dst), _ = model(img)
    loss_old = F.cross_entropy(cls, lbl) + F.cross_entropy(dst, lbl)
    model.zero_grad(set_to_none=True)
    (cls2, dst2), _ = model(img)
    (F.cross_entropy(cls2, lbl) + F.cross_entropy(dst2, lbl)).backward()
    opt.step()
    model(img)
    with pytest.raises(RuntimeError, match="step prologue"):
        loss_old.backward()


@pytest.mark.parametrize("B,K,Nout", [(16, 384, 1000), (128, 192, 1000), (3, 64, 24)])
def test_head_int8_path_matches_torch_composition(Q, B, K, Nout):
    """LSQ_QLinear4head (8-bit classifier head): the int8 tensor-core path (functional.HeadLinearFn) against the op-for-op torch
    composition of qlinear.py:193-238 it replaces (which the tiny-DeiT goldens pin to the reference): output <= 1e-5, every
    gradient (input, both shifts, the scalar input step, weight, per-row weight steps, bias) <= 1e-3."""
    from ofq_b200 import ops
    from ofq_b200.quantization.modules import qlinear as QL
    torch.manual_seed(46)
    lin = nn.Linear(K, Nout)
    mod = QL.LSQ_QLinear4head(m=lin, weight_quant_method="lsq", pretrained_initialized=True).cuda()
    with torch.no_grad():
        mod.move_b4.bias.normal_(0, 0.05)
        mod.move_aft.bias.normal_(0, 0.05)
    x0 = torch.randn(B, K, device="cuda")
    go = torch.randn(B, Nout, device="cuda")
    with torch.no_grad():
        mod(x0)                                        # creates both step sizes (torch composition)
    assert mod._native(x0)
    res = []
    for native in (True, False):
        mod.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        l0 = ops.LAUNCHES
        if native:
            y = mod(x)
        else:
            w = mod.lsqw_fn(mod.weight)
            y = F.linear(mod.move_aft(mod.input_quant_fn(mod.move_b4(x))), w) + mod.bias
        n = ops.LAUNCHES - l0
        y.backward(go)
        res.append((y.detach(), x.grad.clone(), {k: p.grad.clone() for k, p in mod.named_parameters()}, n))
    assert res[0][3] > 0 and res[1][3] == 0                        # the native path really ran the C-ABI kernels
    assert rel_err(res[0][0], res[1][0]) < 1e-5
    assert rel_err(res[0][1], res[1][1]) < 1e-3
    # the scalar input step's gradient is ONE random-walk sum g * sum dx_hat * (q - v) over B*K elements: its error budget is 1e-3
    # of the walk's own scale g * ||dx||_2 (the sum itself can come out arbitrarily close to zero)
    walk = res[1][1].norm().item() / (127.0 * B * K) ** 0.5
    for k, b in res[1][2].items():
        a = res[0][2][k]
        floor = 1e-3 * walk if k == "input_quant_fn.s" else 1e-6 * max(1.0, b.abs().max().item())
        assert rel_err(a, b) < 1e-3 or (a - b).abs().max().item() <= floor, (k, rel_err(a, b))


def test_flat_gradient_buffer_direct_slots_and_zero_arena(Q):
    """ddp.FlatGradAllReduce(direct=True): the dW GEMMs of the quantized layers accumulate straight into the flat gradient
    buffer (autograd adopts the slice as .grad: no gather copy, no per-weight zero fill), and the tiny scratch vectors of the
    backward come from the per-step zero arena. Pure re-plumbing: every gradient equals the plain path's."""
    from ofq_b200 import ops
    from ofq_b200.ddp import FlatGradAllReduce
    from ofq_b200.host.deit import DistilledVisionTransformer
    from ofq_b200.quantization import functional as Fn
    torch.manual_seed(44)
    depth = 2
    model = DistilledVisionTransformer(embed_dim=128, depth=depth, num_heads=2, num_classes=10)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(Q.deit_qmodule_names(depth), 2, 2),
                                             pretrained_initialized=True, qk_reparam=True).cuda()
    img = torch.randn(3, 3, 224, 224, device="cuda")
    lbl = torch.tensor([1, 5, 7], device="cuda")
    model.eval()
    with torch.no_grad():
        model(img)
    model.train()

    def fwd_bwd():
        (cls, dst), _ = model(img)
        (F.cross_entropy(cls, lbl) + F.cross_entropy(dst, lbl)).backward()

    model.zero_grad(set_to_none=True)
    fwd_bwd()
    ref = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    arena = ops._ARENA[img.device]
    assert arena.carved > 0                                       # the backward scratch came from the arena ...
    ddp = FlatGradAllReduce(model.parameters(), 1)
    for _ in range(2):                                            # twice: the second step re-uses the slots after zero()
        ddp.zero()
        assert Fn.GRAD_SLOTS is ddp
        hits0 = ddp.direct_hits
        fwd_bwd()
        weights = [(n, p) for n, p in model.named_parameters() if p.ndim == 2 and ".blocks." in "." + n and n.endswith(".weight")
                   and not n.endswith(("q.weight", "k.weight"))]
        assert ddp.direct_hits - hits0 == len(weights) == 4 * depth          # v, proj, fc1, fc2 of every block
        views = {id(p): v for p, v in zip(ddp.params, ddp.views)}
        for n, p in weights:
            assert p.grad.data_ptr() == views[id(p)].data_ptr(), n           # adopted, not copied
        ddp.reduce()
        assert Fn.GRAD_SLOTS is None
        for n, p in model.named_parameters():
            if n in ref:
                assert p.grad.data_ptr() == views[id(p)].data_ptr()
                # (split-K weight gradients are summed by atomics in arrival order: equal up to fp32 summation order)
                assert torch.equal(p.grad, ref[n]) or rel_err(p.grad, ref[n]) < 1e-5, n


@pytest.mark.parametrize("patch,cout", [(16, 64), (32, 96)])
def test_patch_embed_int8_path_matches_torch_composition(Q, patch, cout):
    """LSQ_QConv2d (8-bit patch embedding): the int8 tensor-core path (functional.PatchEmbedFn) against the op-for-op torch
    composition of qlinear.py:138-191 that it replaces: output <= 1e-5, every gradient (image, per-pixel shifts, the three
    image step sizes, conv weight, weight step sizes, bias) <= 1e-3 (the backward GEMM operand is fp16 range-scaled)."""
    from ofq_b200.quantization.modules import qlinear as QL
    torch.manual_seed(43)
    conv = nn.Conv2d(3, cout, patch, patch)
    mod = QL.LSQ_QConv2d(m=conv, pretrained_initialized=True).cuda()
    with torch.no_grad():
        mod.move_b4.bias.normal_(0, 0.05)
        mod.move_aft.bias.normal_(0, 0.05)
    x0 = torch.randn(4, 3, 224, 224, device="cuda")
    with torch.no_grad():
        mod(x0)                                       # creates the step sizes and latches the sign (torch path)
    gout = torch.randn(4, cout, 224 // patch, 224 // patch, device="cuda")
    res = []
    for fused in (True, False):
        x = x0.clone().requires_grad_(True)
        mod.zero_grad(set_to_none=True)
        saved = QL.F16_BWD
        QL.F16_BWD = saved and fused                  # the switch that routes LSQ_QConv2d.forward to PatchEmbedFn
        try:
            out = mod(x)
        finally:
            QL.F16_BWD = saved
        assert (type(out.grad_fn).__name__ == "PatchEmbedFnBackward") == fused
        (out * gout).sum().backward()
        res.append((out.detach(), x.grad.clone(), {n: p.grad.clone() for n, p in mod.named_parameters()}))
    (o1, dx1, g1), (o0, dx0, g0) = res
    assert rel_err(o1, o0) < 1e-5
    assert rel_err(dx1, dx0) < GRAD_TOL
    assert set(g1) == set(g0)
    for n in g0:
        assert rel_err(g1[n], g0[n]) < GRAD_TOL, n
