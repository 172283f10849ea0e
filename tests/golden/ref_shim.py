"""Import shim that lets the UNMODIFIED reference (nbasyl/OFQ at /root/reference) run on CPU here.

Only used by tests/golden/make_golden.py (build container, where /root/reference is mounted) to produce
the committed golden vectors. Nothing in tests/, bench.py or the package imports the reference at run time.

What is stubbed: the stdlib `imp` and `turtle` modules (gone / no tkinter), the handful of
`timm` names the model files import, and — because the reference hard-codes device="cuda" when it lazily
creates LSQ scales (lsq.py:556-568) — `Tensor.cuda` and `torch.zeros(device="cuda")` on a CPU-only box.
"""
from __future__ import annotations

import collections.abc
import sys
import types
from itertools import repeat

import torch
import torch.nn as nn

REF_ROOT = "/root/reference"


def _to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return tuple(repeat(x, 2))


class _PatchEmbed(nn.Module):
    """timm 0.5.4 PatchEmbed semantics: conv(patch, stride=patch) -> flatten(2).transpose(1, 2) -> Identity."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
        super().__init__()
        img_size = _to_2tuple(img_size)
        patch_size = _to_2tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class _SoftTargetCrossEntropy(nn.Module):
    def forward(self, x, target):
        return torch.sum(-target * torch.nn.functional.log_softmax(x, dim=-1), dim=-1).mean()


def install() -> None:
    if "src" in sys.modules:
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("imp")
    mod("turtle", forward=None)
    timm = mod("timm")
    timm.data = mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225),
                    IMAGENET_INCEPTION_MEAN=(0.5, 0.5, 0.5), IMAGENET_INCEPTION_STD=(0.5, 0.5, 0.5))
    timm.models = mod("timm.models")
    timm.models.helpers = mod("timm.models.helpers", build_model_with_cfg=None, named_apply=None, adapt_input_conv=None)
    timm.models.layers = mod("timm.models.layers", PatchEmbed=_PatchEmbed, DropPath=_DropPath,
                             trunc_normal_=torch.nn.init.trunc_normal_, lecun_normal_=None, to_2tuple=_to_2tuple)
    timm.models.registry = mod("timm.models.registry", register_model=lambda f: f)
    timm.loss = mod("timm.loss", SoftTargetCrossEntropy=_SoftTargetCrossEntropy)

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
        _zeros = torch.zeros

        def zeros(*a, **k):
            if k.get("device") == "cuda":
                k.pop("device")
            return _zeros(*a, **k)

        torch.zeros = zeros
    sys.path.insert(0, REF_ROOT)


def qconfigs(names, wbits, abits):
    """The dict train.py:399-417 builds for `--wq-mode statsq --aq-mode lsq --aq-per-channel --wq-per-channel`."""
    out = {}
    for n in names:
        out[n] = {
            "weight": {"mode": "statsq", "bit": wbits, "all_positive": False, "symmetric": True,
                       "per_channel": True, "normalize_first": False, "learnable": False},
            "act": {"enable": True, "mode": "lsq", "bit": abits, "per_channel": True,
                    "normalize_first": False, "learnable": True},
            "q_attn_dropout": False,
            "act_layer": nn.GELU,
        }
    return out


def deit_qmodule_names(depth=12):
    names = ["patch_embed.proj"]
    for i in range(depth):
        names += [f"blocks.{i}.attn", f"blocks.{i}.mlp"]
    return names + ["head", "head_dist"]
