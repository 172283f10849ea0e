"""Module-swap registry: the drop-in boundary (reference: src/quantization/modules/utils.py:21-413).

`replace_module_by_qmodule_deit(model, qconfigs, pretrained_initialized, qk_reparam, qk_reparam_type,
boundaryRange)` keeps the reference signature: every name in `qconfigs` is looked up in the model and replaced in
place by the quantized class registered for its type.  The patch-embed conv and the heads are always 8/8-bit
LSQ (utils.py:126-156) whatever the flags say.
"""
from __future__ import annotations

import torch

from ... import prologue

from ...host.deit import Attention as deit_attention
from ...host.deit import Mlp
from ...host.swin import MLP as swin_MLP
from ...host.swin import ShiftedWindowAttention
from .attention import QAttention, QAttention_qkreparam, QAttention_qkreparam_4_cga
from .qlinear import LSQ_QConv2d, LSQ_QLinear4head, QLinear, QMLP
from .swin_attention_and_mlp import (QAttention_swin, QAttention_swin_qkreparam, QAttention_swin_qkreparam_4_cga,
                                     QMLP_swin)

QMODULE_MAPPINGS = {torch.nn.Linear: QLinear, deit_attention: QAttention, Mlp: QMLP}
# 0: QAttention_qkreparam, 1: QAttention_qkreparam_4_cga  (utils.py:27-39)
QMODULE_MAPPINGS_QK_REPARAM = [
    {torch.nn.Linear: QLinear, deit_attention: QAttention_qkreparam, Mlp: QMLP},
    {torch.nn.Linear: QLinear, deit_attention: QAttention_qkreparam_4_cga, Mlp: QMLP},
]


# utils.py:286-303
QMODULE_MAPPINGS_SWIN = {torch.nn.Linear: QLinear, ShiftedWindowAttention: QAttention_swin, swin_MLP: QMLP_swin}
QMODULE_MAPPINGS_QK_REPARAM_SWIN = [
    {torch.nn.Linear: QLinear, ShiftedWindowAttention: QAttention_swin_qkreparam, swin_MLP: QMLP_swin},
    {torch.nn.Linear: QLinear, ShiftedWindowAttention: QAttention_swin_qkreparam_4_cga, swin_MLP: QMLP_swin},
]


def _qclass_for(mapping, module, swin=False):
    """The quantized class for a host module. Exact type first (the reference's registry, utils.py:21-44, is keyed on ITS
    host classes); otherwise by interface, so that the reference's own `src.deit_vision_transformer.Attention / Mlp`,
    `src.swin.ShiftedWindowAttention`, torchvision's `MLP` or timm's layers are swapped exactly like the repo's host classes."""
    cls = mapping.get(type(module))
    if cls is not None:
        return cls
    if isinstance(module, torch.nn.Linear):
        return mapping[torch.nn.Linear]
    attn_key, mlp_key = (ShiftedWindowAttention, swin_MLP) if swin else (deit_attention, Mlp)
    if all(hasattr(module, a) for a in ("qkv", "proj", "num_heads")):
        return mapping[attn_key]
    if all(hasattr(module, a) for a in ("fc1", "fc2")):
        return mapping[mlp_key]
    if swin and isinstance(module, torch.nn.Sequential) and len(module) >= 5 and isinstance(module[0], torch.nn.Linear) \
            and isinstance(module[3], torch.nn.Linear):
        return mapping[mlp_key]                      # torchvision.ops.MLP: Sequential(Linear, act, Dropout, Linear, Dropout)
    raise KeyError(f"no quantized module is registered for {type(module).__module__}.{type(module).__name__}")


def get_module_by_name(model, module_name):
    module = model
    for name in module_name.split("."):
        module = getattr(module, name)
    return module


def set_module_by_name(model, module_name, module):
    names = module_name.split(".")
    parent = get_module_by_name(model, ".".join(names[:-1])) if len(names) > 1 else model
    setattr(parent, names[-1], module)


def _eight_bit_kwargs(cfg, pretrained_initialized):
    return dict(weight_bits=8, input_bits=8, weight_channelwise=True, input_channelwise=True, weight_quant_method="lsq",
                input_quant_method="lsq", aq_learnable=True, wq_learnable=True, act_layer=cfg["act_layer"],
                pretrained_initialized=pretrained_initialized)


def replace_module_by_qmodule_deit(model, qconfigs, pretrained_initialized=False, qk_reparam=False, qk_reparam_type=0,
                                   boundaryRange=0.005):
    first = qconfigs[list(qconfigs.keys())[0]]
    if first["weight"]["mode"] == "lsq" and first["act"]["mode"] == "lsq":
        raise NotImplementedError("the LSQ-weight baseline (LSQ_w_and_act_*) is not used by any OFQ script "
                                  "and is outside the B200 hot path (SURVEY.md §8a)")
    mapping = QMODULE_MAPPINGS_QK_REPARAM[qk_reparam_type] if qk_reparam else QMODULE_MAPPINGS
    for name, cfg in qconfigs.items():
        module = get_module_by_name(model, name)
        if name == "patch_embed.proj":
            qmodule = LSQ_QConv2d(m=module, **_eight_bit_kwargs(cfg, pretrained_initialized))
        elif name in ("head", "head_dist"):
            qmodule = LSQ_QLinear4head(m=module, symmetric=True, **_eight_bit_kwargs(cfg, pretrained_initialized))
        else:
            extra = {"boundaryRange": boundaryRange} if (qk_reparam and qk_reparam_type == 1) else {}
            qmodule = _qclass_for(mapping, module)(
                m=module, weight_bits=cfg["weight"]["bit"], input_bits=cfg["act"]["bit"],
                weight_channelwise=cfg["weight"]["per_channel"], input_channelwise=cfg["act"]["per_channel"],
                weight_quant_method=cfg["weight"]["mode"], input_quant_method=cfg["act"]["mode"],
                aq_learnable=cfg["act"]["learnable"], wq_learnable=cfg["weight"]["learnable"],
                act_layer=cfg["act_layer"], pretrained_initialized=pretrained_initialized, **extra)
        set_module_by_name(model, name, qmodule)
    prologue.install(model)     # weight codes / W_qk / step sizes of all layers in three launches per forward (ofq_b200/prologue.py)
    return model


def replace_module_by_qmodule_swin(model, qconfigs, pretrained_initialized=False, qk_reparam=False, qk_reparam_type=0,
                                   boundaryRange=0.005):
    """utils.py:305-413. `features.0.0` (patch-embed conv) and `head` are always 8/8-bit LSQ; `features.N.reduction`
    (nn.Linear, 4-D input) becomes a QLinear whose step sizes are per W' index."""
    mapping = QMODULE_MAPPINGS_QK_REPARAM_SWIN[qk_reparam_type] if qk_reparam else QMODULE_MAPPINGS_SWIN
    for name, cfg in qconfigs.items():
        module = get_module_by_name(model, name)
        if name == "features.0.0":
            qmodule = LSQ_QConv2d(m=module, **_eight_bit_kwargs(cfg, pretrained_initialized))
        elif name == "head":
            qmodule = LSQ_QLinear4head(m=module, symmetric=True, **_eight_bit_kwargs(cfg, pretrained_initialized))
        else:
            qmodule = _qclass_for(mapping, module, swin=True)(
                m=module, weight_bits=cfg["weight"]["bit"], input_bits=cfg["act"]["bit"],
                weight_channelwise=cfg["weight"]["per_channel"], input_channelwise=cfg["act"]["per_channel"],
                weight_quant_method=cfg["weight"]["mode"], input_quant_method=cfg["act"]["mode"],
                aq_learnable=cfg["act"]["learnable"], wq_learnable=cfg["weight"]["learnable"],
                act_layer=cfg["act_layer"], pretrained_initialized=pretrained_initialized)
        set_module_by_name(model, name, qmodule)
    prologue.install(model)
    return model


def swin_qmodule_names(depths=(2, 2, 6, 2)):
    """configs/swin_t_imagenet.attn_q.yml:44-73."""
    names = ["features.0.0"]
    for i, d in enumerate(depths):
        for j in range(d):
            names += [f"features.{2 * i + 1}.{j}.attn", f"features.{2 * i + 1}.{j}.mlp"]
        if i < len(depths) - 1:
            names.append(f"features.{2 * i + 2}.reduction")
    return names + ["head"]


def make_qconfigs(names, wq_bitw, aq_bitw, act_layer=torch.nn.GELU):
    """The per-module dict train.py:399-417 builds for `--wq-mode statsq --aq-mode lsq --*-per-channel`."""
    out = {}
    for n in names:
        out[n] = {
            "weight": {"mode": "statsq", "bit": wq_bitw, "all_positive": False, "symmetric": True, "per_channel": True,
                       "normalize_first": False, "learnable": False},
            "act": {"enable": True, "mode": "lsq", "bit": aq_bitw, "per_channel": True, "normalize_first": False,
                    "learnable": True},
            "q_attn_dropout": False,
            "act_layer": act_layer,
        }
    return out


def deit_qmodule_names(depth=12):
    """configs/ours_imagenet_recipe.attn_q.yml:47-74."""
    names = ["patch_embed.proj"]
    for i in range(depth):
        names += [f"blocks.{i}.attn", f"blocks.{i}.mlp"]
    return names + ["head", "head_dist"]
