"""Golden vectors for the KD losses (reference src/quantization/utils.py:44-77: KLLossSoft, KDLossSoftandHard), produced by the
UNMODIFIED reference classes on CPU in the build container:

    python tests/golden/make_golden_kd.py        ->  tests/golden/kd_loss.npz
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_shim  # noqa: E402

ref_shim.install()
from src.quantization.utils import KLLossSoft, KDLossSoftandHard  # noqa: E402

torch.manual_seed(77)
out = {}
for tag, (B, K) in {"a": (8, 1000), "b": (5, 37)}.items():
    cls = (torch.randn(B, K) * 3).requires_grad_(True)
    dist = (torch.randn(B, K) * 3).requires_grad_(True)
    teacher = torch.randn(B, K) * 4
    teacher_dist = torch.randn(B, K) * 4
    y = torch.randint(0, K, (B,))
    out.update({f"{tag}.cls": cls, f"{tag}.dist": dist, f"{tag}.teacher": teacher, f"{tag}.teacher_dist": teacher_dist, f"{tag}.y": y})
    # KDLossSoftandHard on the distilled student's (cls, dist) tuple with the teacher's training-mode (cls, dist) tuple
    # (train.py:904-910: `soft_target, _ = teacher(input)` with the teacher left in training mode)
    loss = KDLossSoftandHard()((cls, dist), y, (teacher, teacher_dist))
    loss.backward()
    out.update({f"{tag}.sh_tuple.loss": loss.detach(), f"{tag}.sh_tuple.dcls": cls.grad.clone(), f"{tag}.sh_tuple.ddist": dist.grad.clone()})
    cls.grad = dist.grad = None
    # single-logit student (Swin): both terms on the same output
    loss = KDLossSoftandHard()(cls, y, teacher)
    loss.backward()
    out.update({f"{tag}.sh_single.loss": loss.detach(), f"{tag}.sh_single.dcls": cls.grad.clone()})
    cls.grad = None
    # KLLossSoft alone (kd_hard_and_soft == 0, train.py:898-903), with and without a temperature
    for T in (1.0, 2.5):
        loss = KLLossSoft()((cls, dist), (teacher, teacher_dist), T=T)
        loss.backward()
        out.update({f"{tag}.soft_T{T}.loss": loss.detach(), f"{tag}.soft_T{T}.dcls": cls.grad.clone()})
        cls.grad = None
np.savez_compressed(HERE / "kd_loss.npz", **{k: v.detach().numpy() for k, v in out.items()})
print("kd_loss.npz", (HERE / "kd_loss.npz").stat().st_size // 1024, "KiB", len(out), "arrays")
