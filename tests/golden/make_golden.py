"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (nbasyl/OFQ,
mounted read-only at /root/reference) on CPU in the build container.

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so these files are the pin for
oracle/ofq_oracle.py: tests/test_oracle_golden.py replays them without needing /root/reference.
Everything is seeded; inputs that are cheap to regenerate from the seed are still stored so that the
fixtures do not depend on RNG stability across torch versions.
"""
from __future__ import annotations

import ast
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_shim  # noqa: E402

ref_shim.install()
import src  # noqa: E402  (the reference package)
from src.quantization.quantizer.statsq import StatsQuantizer, StatsQuantizer_specific_4_qkreparam_cga  # noqa: E402
from src.quantization.quantizer.lsq import LsqQuantizer, LsqQuantizer4v  # noqa: E402
from src.quantization.modules.qlinear import QLinear, QMLP, LSQ_input  # noqa: E402
from src.quantization.modules.attention import QAttention, QAttention_qkreparam, QAttention_qkreparam_4_cga  # noqa: E402
from src.quantization.modules.utils import replace_module_by_qmodule_deit  # noqa: E402
from src.deit_vision_transformer import Attention, Mlp  # noqa: E402
from src.deit import DistilledVisionTransformer  # noqa: E402
from functools import partial  # noqa: E402


def npify(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def save(name, d):
    path = HERE / f"{name}.npz"
    np.savez_compressed(path, **npify(d))
    print(f"{path.name}: {path.stat().st_size / 1024:.0f} KiB, {len(d)} arrays")


def grads_of(module, prefix="grad."):
    return {prefix + n: p.grad for n, p in module.named_parameters() if p.grad is not None}


# --------------------------------------------------------------------------------------- quantizers
def gen_statsq():
    out = {}
    kat = torch.tensor([[.10, -.20, .30, -.40, .05, 0., 1., -1.]])
    torch.manual_seed(1)
    w = torch.randn(48, 40) * 0.02
    dy = torch.randn(64, 72) / 8                      # dyadic-ish: sums exact in any order
    dyadic = torch.round(dy * 1024) / 1024
    out["kat.w"], out["rand.w"], out["dyadic.w"] = kat, w, dyadic
    for bits in (2, 3, 4):
        for name, t in (("kat", kat), ("rand", w), ("dyadic", dyadic)):
            q = StatsQuantizer(num_bits=bits, clip_learnable=False)
            tt = t.clone().requires_grad_(True)
            y = q(tt)
            y.backward(torch.ones_like(y) * 0.5)
            out[f"{name}.b{bits}.out"] = y
            out[f"{name}.b{bits}.grad"] = tt.grad
            out[f"{name}.b{bits}.s"] = q.s
            q2 = StatsQuantizer_specific_4_qkreparam_cga(num_bits=bits, clip_learnable=False, boundaryRange=0.005)
            q2.train()
            out[f"{name}.b{bits}.out_cga"] = q2(t.clone())
    save("statsq", out)


def gen_lsq():
    out = {}
    torch.manual_seed(2)
    x3 = torch.randn(3, 10, 24)
    x4 = torch.rand(2, 3, 10, 10).softmax(-1)
    out["x3"], out["x4"] = x3, x4
    for bit in (2, 3, 4):
        for pos in (False, True):
            tag = f"b{bit}.{'u' if pos else 's'}"
            # per-token quantizer on a 3-D activation, lazily initialised by the first forward
            q = LsqQuantizer(bit=bit, all_positive=pos, per_channel=True, learnable=True)
            xin = (x3.abs() if pos else x3).clone().requires_grad_(True)
            y = q(xin)
            out[f"rows3.{tag}.s_init"] = q.s.detach().clone()
            # perturb the scale so that clipping and the floor path are exercised, then a real fwd+bwd
            with torch.no_grad():
                q.s.mul_(torch.linspace(0.3, 1.7, q.s.numel()))
                q.s[0] = 1e-7
            y = q(xin)
            go = torch.randn_like(y)
            y.backward(go)
            out[f"rows3.{tag}.s"], out[f"rows3.{tag}.out"], out[f"rows3.{tag}.go"] = q.s.detach().clone(), y, go
            out[f"rows3.{tag}.dx"], out[f"rows3.{tag}.ds"] = xin.grad, q.s.grad
            # per-channel (4v) quantizer
            q = LsqQuantizer4v(bit=bit, all_positive=pos, per_channel=True, learnable=True)
            xin = (x3.abs() if pos else x3).clone().requires_grad_(True)
            q(xin)
            out[f"cols3.{tag}.s_init"] = q.s.detach().clone()
            with torch.no_grad():
                q.s.mul_(torch.linspace(0.5, 1.5, q.s.numel()))
            y = q(xin)
            y.backward(go)
            out[f"cols3.{tag}.s"], out[f"cols3.{tag}.out"] = q.s.detach().clone(), y
            out[f"cols3.{tag}.dx"], out[f"cols3.{tag}.ds"] = xin.grad, q.s.grad
        # probabilities: 4-D, unsigned, scale per query row
        q = LsqQuantizer(bit=bit, all_positive=True, per_channel=True, learnable=True)
        xin = x4.clone().requires_grad_(True)
        y = q(xin)
        go4 = torch.randn_like(y)
        y.backward(go4)
        out[f"rows4.b{bit}.s"], out[f"rows4.b{bit}.out"], out[f"rows4.b{bit}.go"] = q.s.detach().clone(), y, go4
        out[f"rows4.b{bit}.dx"], out[f"rows4.b{bit}.ds"] = xin.grad, q.s.grad
    save("lsq", out)


# --------------------------------------------------------------------------------------- layers
def run_module(mod, x, init_first=True):
    """setup_alpha semantics (train.py:997-1010): one no-grad forward creates the scales, then fwd+bwd."""
    if init_first:
        mod.eval()
        with torch.no_grad():
            mod(x)
    mod.train()
    xin = x.clone().requires_grad_(True)
    y = mod(xin)
    y = y[0] if isinstance(y, tuple) else y
    go = torch.randn(y.shape, generator=torch.Generator().manual_seed(99))
    y.backward(go)
    d = {"x": x, "out": y, "go": go, "dx": xin.grad}
    d.update({"param." + k: v for k, v in mod.state_dict().items()})
    d.update(grads_of(mod))
    return d


def randomize_shifts(mod, std=0.05, seed=7):
    """The learnable shifts start at zero; give them values so the affine-offset terms are exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if "move_" in n and n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * std)


def gen_layers():
    torch.manual_seed(3)
    B, N, C, H = 2, 10, 32, 2
    for bits in (2, 4):
        lin = nn.Linear(C, 48)
        q = QLinear(m=lin, weight_bits=bits, input_bits=bits, pretrained_initialized=True)
        randomize_shifts(q)
        save(f"qlinear_w{bits}a{bits}", run_module(q, torch.randn(B, N, C)))

        mlp = Mlp(in_features=C, hidden_features=4 * C)
        q = QMLP(m=mlp, weight_bits=bits, input_bits=bits, act_layer=nn.GELU, pretrained_initialized=True)
        randomize_shifts(q)
        save(f"qmlp_w{bits}a{bits}", run_module(q, torch.randn(B, N, C)))

        attn = Attention(C, num_heads=H, qkv_bias=True)
        q = QAttention(attn, weight_bits=bits, input_bits=bits, pretrained_initialized=True)
        randomize_shifts(q)
        save(f"qattention_w{bits}a{bits}", run_module(q, torch.randn(B, N, C)))

        attn = Attention(C, num_heads=H, qkv_bias=True)
        q = QAttention_qkreparam(attn, weight_bits=bits, input_bits=bits, pretrained_initialized=True)
        randomize_shifts(q)
        xq = torch.randn(B, N, C)
        d = run_module(q, xq)
        # the _4_cga variant must be value- and gradient-identical
        q2 = QAttention_qkreparam_4_cga(attn, weight_bits=bits, input_bits=bits, pretrained_initialized=True,
                                        boundaryRange=0.005)
        q2.load_state_dict({k: v for k, v in q.state_dict().items() if not k.endswith(".s")}, strict=False)
        d2 = run_module(q2, xq)
        d["cga_variant_identical"] = torch.tensor(
            float(torch.equal(d["out"], d2["out"]) and torch.equal(d["dx"], d2["dx"])
                  and all(torch.equal(d[k], d2[k]) for k in d if k.startswith("grad."))))
        assert d["cga_variant_identical"].item() == 1.0
        save(f"qattention_qkr_w{bits}a{bits}", d)


def gen_deit():
    """A depth-2, width-64 distilled DeiT with every qmodule of configs/ours_imagenet_recipe.attn_q.yml quantized
    (image size must stay 224: LearnableBias4img(224*224) is hard-coded, qlinear.py:161-162)."""
    for qkr in (False, True):
        torch.manual_seed(4)
        model = DistilledVisionTransformer(img_size=224, patch_size=16, embed_dim=64, depth=2, num_heads=2,
                                           mlp_ratio=4, qkv_bias=True, num_classes=10,
                                           norm_layer=partial(nn.LayerNorm, eps=1e-6), act_layer=nn.GELU)
        names = ref_shim.deit_qmodule_names(depth=2)
        model = replace_module_by_qmodule_deit(model, ref_shim.qconfigs(names, 2, 2), pretrained_initialized=True,
                                               qk_reparam=qkr, qk_reparam_type=0)
        randomize_shifts(model, std=0.02)
        img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(5))
        labels = torch.tensor([3, 7])
        model.eval()
        with torch.no_grad():
            ev0 = model(img)[0]
        model.train()
        (cls, dist), _ = model(img)
        loss = nn.functional.cross_entropy(cls, labels) + nn.functional.cross_entropy(dist, labels)
        loss.backward()
        model.eval()
        with torch.no_grad():
            ev = model(img)[0]
        d = {"img_seed": 5, "labels": labels, "cls": cls, "dist": dist, "loss": loss, "eval_logits": ev,
             "eval_logits_first": ev0, "signed": model.patch_embed.proj.input_quant_fn.signed}
        sd = model.state_dict()
        d.update({"param." + k: v for k, v in sd.items()})
        # gradients: keep the small ones whole, summarise big ones by slices + norms
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            g = p.grad
            d["gnorm." + n] = g.norm()
            d["grad." + n] = g if g.numel() <= 4096 else g.flatten()[:: max(1, g.numel() // 2048)][:2048]
        save(f"deit_tiny2_{'qkr' if qkr else 'plain'}_w2a2", d)


# --------------------------------------------------------------------------------------- Swin
def gen_swin():
    from src.swin import ShiftedWindowAttention, SwinTransformer
    from src.quantization.modules.swin_attention_and_mlp import QAttention_swin, QAttention_swin_qkreparam
    from src.quantization.modules.utils import replace_module_by_qmodule_swin
    torch.manual_seed(9)
    dim, heads = 32, 2
    for shift in (0, 3):
        for cls, tag in ((QAttention_swin, "plain"), (QAttention_swin_qkreparam, "qkr")):
            m = ShiftedWindowAttention(dim, [7, 7], [shift, shift], heads)
            q = cls(m, weight_bits=3, input_bits=3, pretrained_initialized=True)
            randomize_shifts(q)
            with torch.no_grad():
                q.relative_position_bias_table.copy_(torch.randn_like(q.relative_position_bias_table) * 0.2)
            d = run_module(q, torch.randn(2, 14, 14, dim))
            d = {k: v for k, v in d.items() if not k.endswith("relative_position_index")}
            save(f"qattention_swin_{tag}_shift{shift}_w3a3", d)
    for qkr in (False, True):
        torch.manual_seed(10)
        model = SwinTransformer(patch_size=[4, 4], embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=[7, 7], num_classes=10)
        names = ["features.0.0", "features.1.0.attn", "features.1.0.mlp", "features.1.1.attn", "features.1.1.mlp",
                 "features.2.reduction", "features.3.0.attn", "features.3.0.mlp", "features.3.1.attn", "features.3.1.mlp", "head"]
        model = replace_module_by_qmodule_swin(model, ref_shim.qconfigs(names, 3, 3), pretrained_initialized=True,
                                               qk_reparam=qkr, qk_reparam_type=0)
        randomize_shifts(model, std=0.02)
        img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(11))
        labels = torch.tensor([2, 8])
        model.eval()
        with torch.no_grad():
            model(img)
        model.train()
        logits, _ = model(img)
        loss = nn.functional.cross_entropy(logits, labels)
        loss.backward()
        d = {"img_seed": 11, "labels": labels, "logits": logits, "loss": loss}
        d.update({"param." + k: v for k, v in model.state_dict().items() if not k.endswith("relative_position_index")})
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            g = p.grad
            d["gnorm." + n] = g.norm()
            d["grad." + n] = g if g.numel() <= 4096 else g.flatten()[:: max(1, g.numel() // 2048)][:2048]
        save(f"swin_tiny2_{'qkr' if qkr else 'plain'}_w3a3", d)


# --------------------------------------------------------------------------------------- CGA
def load_reference_function(path, name):
    """exec a single top-level function of a reference script without importing the script (cga.py needs timm)."""
    tree = ast.parse(Path(path).read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def gen_cga():
    freeze = load_reference_function(os.path.join(ref_shim.REF_ROOT, "cga.py"), "freeze_outside_boundary_weight_idx")
    out = {}
    torch.manual_seed(6)
    w = torch.nn.init.trunc_normal_(torch.empty(96, 64), std=0.02)
    # plant weights exactly on / next to rounding boundaries of row 0 so the band edges are exercised
    out["w"] = w
    for bits in (2, 3, 4):
        for br in (0.005, 0.05):
            out[f"mask.b{bits}.br{br}"] = freeze(w, bits, boundaryRange=br)
    # one masked AdamW step exactly as cga.py:953-1013 does it for one module, with torch.optim.AdamW
    lin = nn.Linear(64, 96, bias=False)
    with torch.no_grad():
        lin.weight.copy_(w)
    opt = torch.optim.AdamW(lin.parameters(), lr=1e-3, weight_decay=0.05)
    gen = torch.Generator().manual_seed(8)
    for step in range(3):
        g = torch.randn(96, 64, generator=gen) * 1e-3
        lin.weight.grad = g.clone()
        out[f"step{step}.grad"] = g
        f = freeze(lin.weight, 2, boundaryRange=0.05)
        lin.weight.grad = lin.weight.grad * f * 0.0 + lin.weight.grad * (1 - f)
        stash = (lin.weight * f).detach().clone()
        opt.step()
        with torch.no_grad():
            keep = lin.weight.detach().clone() * (1 - f)
            lin.weight.data.copy_(keep + stash)
        st = opt.state[lin.weight]
        out[f"step{step}.w"], out[f"step{step}.mask"] = lin.weight.detach().clone(), f
        out[f"step{step}.exp_avg"], out[f"step{step}.exp_avg_sq"] = st["exp_avg"].clone(), st["exp_avg_sq"].clone()
    save("cga", out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    gen_statsq()
    gen_lsq()
    gen_layers()
    gen_deit()
    gen_swin()
    gen_cga()
