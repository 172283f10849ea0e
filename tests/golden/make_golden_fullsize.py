"""Full-size golden vectors: the UNMODIFIED reference (nbasyl/OFQ at /root/reference, CPU, fp32) at the real width and
depth of every BASELINE.json configuration, batch 8, forward + backward.

    python tests/golden/make_golden_fullsize.py [config ...]

Writes tests/golden/full_<config>.npz. Parameters and images are regenerated from their names by
tests/fullsize_common.py (bit-identical on any platform), so a fixture holds only: the data-derived LSQ step sizes the
reference's first forward created (`setup_alpha`, train.py:997-1010), logits / loss / eval logits, a strided sample and
the norm of every block output and of every gradient, and, for the first and the last block (first two images), every
integer code tensor of the block (activation codes of all quantizers, StatsQ weight codes).
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
import ref_shim  # noqa: E402

ref_shim.install()
import src  # noqa: E402,F401  (the reference package)
from src.deit import deit_small_distilled_patch16_224, deit_tiny_distilled_patch16_224  # noqa: E402
from src.quantization.modules.utils import replace_module_by_qmodule_deit, replace_module_by_qmodule_swin  # noqa: E402
from src.quantization.quantizer import lsq as ref_lsq  # noqa: E402
from src.swin import swin_t  # noqa: E402

import fullsize_common as FC  # noqa: E402


def build(cfg_name):
    model_name, wb, ab, qkr, qkr_type, _ = FC.CONFIGS[cfg_name]
    torch.manual_seed(0)
    if model_name == "swin_tiny":
        model = swin_t(drop_path=0.0, num_classes=1000)
        model = replace_module_by_qmodule_swin(model, ref_shim.qconfigs(FC.swin_names(), wb, ab), pretrained_initialized=True,
                                               qk_reparam=qkr, qk_reparam_type=qkr_type)
    else:
        ctor = deit_small_distilled_patch16_224 if model_name == "deit_small" else deit_tiny_distilled_patch16_224
        model = ctor(num_classes=1000)
        model = replace_module_by_qmodule_deit(model, ref_shim.qconfigs(FC.deit_names(12), wb, ab), pretrained_initialized=True,
                                               qk_reparam=qkr, qk_reparam_type=qkr_type)
    return model, model_name


def block_names(model_name):
    """(first, last) block module names and the list of all blocks whose outputs are sampled."""
    if model_name == "swin_tiny":
        blocks = [f"features.{2 * i + 1}.{j}" for i, d in enumerate((2, 2, 6, 2)) for j in range(d)]
    else:
        blocks = [f"blocks.{i}" for i in range(12)]
    return blocks


def lsq_codes(mod, x_in, out):
    """Integer codes of a reference LSQ quantizer call: out = round(clamp(x / s')) * s'."""
    with torch.no_grad():
        eps = torch.tensor(1e-5).float()
        name = type(mod).__name__
        if name == "LsqQuantizer4v":
            alpha = mod.s
            g = 1.0 / ((mod.thd_pos * x_in.numel() // x_in.shape[-1]) ** 0.5)
        else:
            alpha = mod.s.unsqueeze(-1)
            g = 1.0 / ((mod.thd_pos * (x_in.numel() // x_in.shape[-2])) ** 0.5)
        se = ref_lsq.grad_scale(ref_lsq.clip(alpha, eps), g)
        q = torch.round(out / se)
        assert torch.equal(q * se, out), "code extraction must reproduce the reference output exactly"
        return q.to(torch.int8)


def statsq_codes(mod, w, out):
    """Odd integer codes 2k+1 of a reference StatsQuantizer call: out = sf * (k + 0.5) / n (up to the STE add-subtract)."""
    with torch.no_grad():
        n = 2 ** (mod.num_bits - 1)
        sf = 2 * w.detach().abs().mean(dim=1, keepdim=True)
        k2 = torch.round(out.detach() / sf * (2 * n))
        assert bool((k2 % 2 != 0).all()) and int(k2.abs().max()) <= 2 ** mod.num_bits - 1
        return k2.to(torch.int8)


def run(cfg_name):
    t0 = time.time()
    model, model_name = build(cfg_name)
    values, chk = FC.fill_state(model.state_dict())
    FC.apply_state(model, values)
    img = FC.det_images()
    labels = FC.det_labels(FC.BATCH, 1000)
    swin = model_name == "swin_tiny"
    # setup_alpha (train.py:997-1010): an eval no-grad forward creates every LSQ step size from this batch
    model.eval()
    with torch.no_grad():
        ev0 = model(img)[0]
    d = {"checksum": np.int64(chk), "batch": FC.BATCH, "eval_logits_first": ev0}
    sd = model.state_dict()
    for k, v in sd.items():
        if not FC.is_regenerated(k, v) and v.is_floating_point():
            d["state." + k] = v.clone()
    if not swin:
        d["signed"] = model.patch_embed.proj.input_quant_fn.signed
    else:
        d["signed"] = model.features[0][0].input_quant_fn.signed

    blocks = block_names(model_name)
    mods = dict(model.named_modules())
    hooks = []
    for bi, bn in enumerate(blocks):
        def out_hook(m, inp, out, bi=bi):
            o = out[0] if isinstance(out, tuple) else out
            d[f"block{bi}.out_sample"] = FC.sample(o, 4096)
            d[f"block{bi}.out_norm"] = o.detach().double().norm().float()
        hooks.append(mods[bn].register_forward_hook(out_hook))
    nimg = FC.CONFIGS[cfg_name][5]
    d["code_images"] = nimg
    for tag, bn in ((("first", blocks[0]), ("last", blocks[-1])) if nimg > 0 else ()):
        def in_hook(m, inp, tag=tag):
            x = inp[0][0] if isinstance(inp[0], tuple) else inp[0]        # the reference's Swin blocks pass (x, attn) tuples
            d[f"{tag}.block_in"] = x.detach()[:nimg].clone()
        hooks.append(mods[bn].register_forward_pre_hook(in_hook))
        for name, m in model.named_modules():
            if not name.startswith(bn + "."):
                continue
            rel = name[len(bn) + 1:]
            cls = type(m).__name__
            if cls.startswith("LsqQuantizer"):
                def qhook(m, inp, out, key=f"{tag}.codes.{rel}"):
                    c = lsq_codes(m, inp[0], out)
                    # activations: first images only; windows of the first images for Swin (batch dim = B * nW)
                    per = c.shape[0] // FC.BATCH
                    d[key] = c[: nimg * per].clone()
                hooks.append(m.register_forward_hook(qhook))
            elif cls.startswith("StatsQuantizer"):
                def whook(m, inp, out, key=f"{tag}.wcodes.{rel}"):
                    d[key] = FC.row_sample(statsq_codes(m, inp[0], out))
                hooks.append(m.register_forward_hook(whook))

    model.train()
    out = model(img)[0]
    if swin:
        loss = F.cross_entropy(out, labels)
        d["logits"] = out
    else:
        cls_l, dist_l = out
        loss = F.cross_entropy(cls_l, labels) + F.cross_entropy(dist_l, labels)
        d["cls"], d["dist"] = cls_l, dist_l
    d["loss"] = loss
    loss.backward()
    for h in hooks:
        h.remove()
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        d["gnorm." + n] = p.grad.double().norm().float()
        d["grad." + n] = FC.sample(p.grad, 2048)
    model.eval()
    with torch.no_grad():
        d["eval_logits"] = model(img)[0]
    path = HERE / f"full_{cfg_name}.npz"
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()})
    print(f"{path.name}: {path.stat().st_size / 1e6:.2f} MB, {len(d)} arrays, loss {loss.item():.6f}, {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(8)
    for name in (sys.argv[1:] or list(FC.CONFIGS)):
        run(name)
