"""Shared by tests/golden/make_golden_fullsize.py (runs the UNMODIFIED reference on CPU, build container only) and by the
full-size parity tests: the BASELINE.json configurations at their real width / depth, and a platform-independent
deterministic parameter / image generator.

A depth-12 DeiT-S has 22.7 M parameters (91 MB): far too much to commit, so the fixtures store only what is data derived
(the lazily created LSQ step sizes) and every other tensor is REGENERATED from its state-dict key by `det_tensor`:
a counter-based splitmix64 stream (pure 64-bit integer arithmetic in numpy), four 16-bit uniforms summed per value
(Irwin-Hall, near-normal), one exact int -> fp32 conversion and one IEEE multiply. No libm call, no torch RNG: the same
bits on any CPU, any thread count and any library version. A checksum of everything generated is stored in the fixture.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch

# name -> (model, wbits, abits, qk_reparam, qk_reparam_type, images whose code tensors are stored)
# BASELINE.json configs in brackets
CONFIGS = {
    "deit_tiny_plain_w4a4": ("deit_tiny", 4, 4, False, 0, 2),      # [1] DeiT-T W4A4 plain quantized attention
    "deit_tiny_qkr_w2a2": ("deit_tiny", 2, 2, True, 0, 2),         # [2] DeiT-T W2A2 attn_q + QKR
    "deit_small_qkr_w2a2": ("deit_small", 2, 2, True, 0, 2),       # [3] DeiT-S W2A2 attn_q (QKR, the headline)
    "deit_small_qkr1_w2a2": ("deit_small", 2, 2, True, 1, 0),      # [5] DeiT-S W2A2, qk_reparam_type 1 (the CGA model;
                                                                   #     value-identical to type 0: no code tensors)
    "swin_tiny_plain_w3a3": ("swin_tiny", 3, 3, False, 0, 1),      # [4] Swin-T W3A3 quantized shifted-window attention
    "swin_tiny_qkr_w3a3": ("swin_tiny", 3, 3, True, 0, 1),         # [4] ... with QKR
}
MODEL_DIMS = {"deit_tiny": dict(embed_dim=192, depth=12, num_heads=3), "deit_small": dict(embed_dim=384, depth=12, num_heads=6)}
BATCH = 8
IMG_SEED, LABEL_SEED = 7001, 7002

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(idx: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(seed & 0xFFFFFFFFFFFFFFFF) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def det_normal(shape, seed: int, std: float = 1.0, mean: float = 0.0) -> torch.Tensor:
    """Near-normal fp32 tensor (sum of four 16-bit uniforms), bit-identical on every platform."""
    n = int(np.prod(shape)) if len(shape) else 1
    r = _splitmix64(np.arange(n, dtype=np.uint64), seed)
    s = ((r & np.uint64(0xFFFF)) + ((r >> np.uint64(16)) & np.uint64(0xFFFF)) + ((r >> np.uint64(32)) & np.uint64(0xFFFF))
         + (r >> np.uint64(48))).astype(np.int64) - 2 * 65535          # in [-131070, 131070], variance 4 * (65536^2 - 1) / 12
    unit = np.float32(1.0 / np.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0))
    v = s.astype(np.float32) * np.float32(np.float32(std) * unit)       # one exact conversion, one IEEE multiply
    if mean != 0.0:
        v = v + np.float32(mean)
    return torch.from_numpy(v.reshape(tuple(shape)))


def det_labels(n: int, classes: int, seed: int = LABEL_SEED) -> torch.Tensor:
    return torch.from_numpy((_splitmix64(np.arange(n, dtype=np.uint64), seed) % np.uint64(classes)).astype(np.int64))


def det_images(batch: int = BATCH) -> torch.Tensor:
    return det_normal((batch, 3, 224, 224), IMG_SEED)


def key_seed(key: str) -> int:
    return zlib.crc32(key.encode()) + 0x5EED0000


def is_regenerated(key: str, t: torch.Tensor) -> bool:
    """Everything is regenerated except data-derived LSQ step sizes (`.s`), the frozen StatsQ `clip_val`, the sign latch of
    the image quantizer and integer buffers."""
    return t.is_floating_point() and not key.endswith((".s", "clip_val", ".signed"))


def det_value(key: str, shape) -> torch.Tensor:
    seed = key_seed(key)
    last = key.rsplit(".", 1)[-1]
    if ".norm" in "." + key or key.startswith("norm"):
        if last == "weight":
            return det_normal(shape, seed, 0.05, 1.0)
        return det_normal(shape, seed, 0.02)
    if last == "relative_position_bias_table":
        return det_normal(shape, seed, 0.1)
    return det_normal(shape, seed, 0.02)          # weights, biases, learnable shifts, tokens, position embedding


def fill_state(sd: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], int]:
    """Deterministic values for every regenerated key of a state dict; returns (new values, checksum)."""
    out, chk = {}, 0
    for k in sorted(sd):
        t = sd[k]
        if not is_regenerated(k, t):
            continue
        v = det_value(k, tuple(t.shape))
        out[k] = v
        chk = (chk * 1000003 + int(v.numpy().view(np.uint32).astype(np.uint64).sum())) % (1 << 61)
    return out, chk


def apply_state(model: torch.nn.Module, values: Dict[str, torch.Tensor]) -> None:
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in values.items():
            sd[k].copy_(v)


def sample(t: torch.Tensor, n: int = 2048) -> torch.Tensor:
    """The strided sample the fixtures keep of a big tensor."""
    f = t.detach().flatten()
    return f if f.numel() <= 2 * n else f[:: max(1, f.numel() // n)][:n].clone()


def row_sample(t: torch.Tensor, n: int = 384) -> torch.Tensor:
    """Weight-code tensors are kept as at most ~n whole rows (StatsQ is row-wise: a row subset is a faithful sample)."""
    r = t.shape[0]
    return t if r <= n else t[:: (r + n - 1) // n].clone()


def swin_names(depths: Iterable[int] = (2, 2, 6, 2)):
    """configs/swin_t_imagenet.attn_q.yml:44-73."""
    depths = tuple(depths)
    names = ["features.0.0"]
    for i, d in enumerate(depths):
        for j in range(d):
            names += [f"features.{2 * i + 1}.{j}.attn", f"features.{2 * i + 1}.{j}.mlp"]
        if i < len(depths) - 1:
            names.append(f"features.{2 * i + 2}.reduction")
    return names + ["head"]


def deit_names(depth: int = 12):
    """configs/ours_imagenet_recipe.attn_q.yml:47-74."""
    names = ["patch_embed.proj"]
    for i in range(depth):
        names += [f"blocks.{i}.attn", f"blocks.{i}.mlp"]
    return names + ["head", "head_dist"]


# ------------------------------------------------------------------------------------------------ model builders
def build_repo_model(cfg_name: str):
    """The repo's host model + drop-in quantized modules for a CONFIGS entry (CPU construction, random init)."""
    import ofq_b200.quantization as Q
    from ofq_b200.host.deit import DistilledVisionTransformer
    model_name, wb, ab, qkr, qkr_type, _ = CONFIGS[cfg_name]
    if model_name == "swin_tiny":
        from ofq_b200.host.swin import swin_t
        model = swin_t(num_classes=1000)
        model = Q.replace_module_by_qmodule_swin(model, Q.make_qconfigs(swin_names(), wb, ab), pretrained_initialized=True,
                                                 qk_reparam=qkr, qk_reparam_type=qkr_type)
    else:
        model = DistilledVisionTransformer(num_classes=1000, **MODEL_DIMS[model_name])
        model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(deit_names(12), wb, ab), pretrained_initialized=True,
                                                 qk_reparam=qkr, qk_reparam_type=qkr_type)
    return model


def golden_state(g: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The stored (data-derived) part of the reference state dict: LSQ step sizes and clip_val."""
    return {k[len("state."):]: v for k, v in g.items() if k.startswith("state.")}


def load_repo_model(cfg_name: str, g: Dict[str, torch.Tensor]):
    """Repo model with the regenerated parameters and the fixture's step sizes; asserts the regeneration checksum."""
    model = build_repo_model(cfg_name)
    values, chk = fill_state(model.state_dict())
    assert chk == int(g["checksum"]), "regenerated parameters differ from the ones the reference fixture was made with"
    apply_state(model, values)
    missing, unexpected = model.load_state_dict(golden_state(g), strict=False)
    assert not unexpected, unexpected
    return model


def oracle_params(cfg_name: str, g: Dict[str, torch.Tensor], requires_grad: bool = True) -> Dict[str, torch.Tensor]:
    """Flat parameter dict (reference key names) for oracle/ofq_oracle.py: regenerated values + the fixture's step sizes."""
    model = build_repo_model(cfg_name)
    sd = model.state_dict()
    values, chk = fill_state(sd)
    assert chk == int(g["checksum"]), "regenerated parameters differ from the ones the reference fixture was made with"
    P = {k: v.clone() for k, v in sd.items()}
    P.update(values)
    P.update({k: v.clone() for k, v in golden_state(g).items()})
    if requires_grad:
        P = {k: (v.requires_grad_(True) if v.is_floating_point() and not k.endswith("clip_val") else v) for k, v in P.items()}
    return P


def block_prefixes(model_name: str):
    if model_name == "swin_tiny":
        return [f"features.{2 * i + 1}.{j}." for i, d in enumerate((2, 2, 6, 2)) for j in range(d)]
    return [f"blocks.{i}." for i in range(12)]


def oracle_forward(cfg_name: str, P, img, signed: int):
    """Training-mode forward of the oracle for a CONFIGS entry. Returns the tuple of logits tensors."""
    from oracle import ofq_oracle as O
    model_name, wb, ab, qkr, _, _ = CONFIGS[cfg_name]
    state = {"signed": int(signed)}
    if model_name == "swin_tiny":
        return (O.swin_forward(img, P, (2, 2, 6, 2), (3, 6, 12, 24), wb, ab, qkr, state),)
    dims = MODEL_DIMS[model_name]
    return O.deit_forward(img, P, dims["depth"], dims["num_heads"], wb, ab, qkr, state)


# ------------------------------------------------------------------------------------------------ code comparison
TIE_TOL = 2e-5     # relative (to max(1, |v|)) window in which two pre-round values count as the same value up to GEMM round-off


def compare_codes(mine, v_mine, ref, v_ref=None, lo=None, hi=None):
    """Mismatch statistics of one integer code tensor against a reference one. With the pre-round values v = x / s of both
    sides, a mismatch is a PROVEN TIE when the two values differ by no more than TIE_TOL * max(1, |v|): both are then within
    that distance of the rounding boundary that separates the two codes. With the clamp bounds [lo, hi] also the
    straight-through mask 1[lo <= v <= hi] of the backward is compared ("mask_flips": same code, but the two pre-round values
    lie on different sides of a clamp bound - the tie class of the gradient; lsq.py:595-599)."""
    mine = mine.detach().cpu().to(torch.int16).reshape(ref.shape)
    bad = mine != ref.to(torch.int16)
    n = int(bad.sum())
    r = {"numel": ref.numel(), "mismatches": n}
    if v_ref is not None:
        vm = v_mine.detach().cpu().reshape(ref.shape).float()
        v_ref = v_ref.float()
        d = (vm - v_ref).abs()
        inside = v_ref.abs() < 1e3
        r["max_pre_round_diff"] = float((d[inside] / v_ref[inside].abs().clamp_min(1.0)).max()) if bool(inside.any()) else 0.0
        if n:
            ties = d[bad] <= TIE_TOL * v_ref[bad].abs().clamp_min(1.0)
            r["ties"] = int(ties.sum())
            r["not_ties"] = n - int(ties.sum())
        else:
            r["ties"] = r["not_ties"] = 0
        if lo is not None:
            flip = ((vm >= lo) & (vm <= hi)) != ((v_ref >= lo) & (v_ref <= hi))
            r["mask_flips"] = int(flip.sum())
            r["mask_not_ties"] = int((d[flip] > TIE_TOL * v_ref[flip].abs().clamp_min(1.0)).sum()) if r["mask_flips"] else 0
    return r


def oracle_block(cfg_name: str, P, pre: str, x, bi: int):
    """One host block (pre-norm attention + MLP with residuals) of a CONFIGS entry, evaluated by the oracle."""
    import torch.nn.functional as F
    from oracle import ofq_oracle as O
    model_name, wb, ab, qkr, _, _ = CONFIGS[cfg_name]
    if model_name != "swin_tiny":
        return O.deit_block(x, P, pre, MODEL_DIMS[model_name]["num_heads"], wb, ab, qkr)
    stage = next(i for i, e in enumerate((2, 4, 10, 12)) if bi < e)
    j = bi - (0, 2, 4, 10)[stage]
    heads = (3, 6, 12, 24)[stage]
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), P[pre + "norm1.weight"], P[pre + "norm1.bias"], 1e-5)
    shift = (0, 0) if j % 2 == 0 else (3, 3)
    x = x + O.swin_window_attention(h, P, pre + "attn.", heads, wb, ab, qkr, (7, 7), shift)
    h = F.layer_norm(x, (C,), P[pre + "norm2.weight"], P[pre + "norm2.bias"], 1e-5)
    return x + O.qmlp(h, P, pre + "mlp.", wb, ab)


def lsq_sites(taps: dict, pre: str):
    """[(site name relative to the block, tap)] of the LSQ quantizer calls of one block, in forward order."""
    return [(k[len(pre):], v) for k, v in taps.items() if k.startswith(pre) and "codes" in v]


def weight_of_site(P, pre: str, name: str):
    """The fp32 weight a StatsQ quantizer of a block sees, by the reference's quantizer name (relative to the block)."""
    from oracle import ofq_oracle as O
    if name.endswith("qk_quant"):
        a = pre + name[: -len("qk_quant")]
        heads = P[a + "quan_a_qkx_fn.s"].numel() // P[a + "quant_x_4_qkv.input_quant_fn.s"].numel()
        return O.wqk_compose(P[a + "q.weight"].detach(), P[a + "k.weight"].detach(), heads)
    if name.endswith("v_quant"):
        return P[pre + name[: -len("v_quant")] + "v.weight"].detach()
    return P[pre + name[: -len("statsq_fn")] + "weight"].detach()
