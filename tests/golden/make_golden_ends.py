"""Golden vectors for the 8-bit ends (reference src/quantization/modules/qlinear.py:138-252: LSQ_QConv2d patch embedding,
LSQ_QLinear4head classifier head), produced by the UNMODIFIED reference classes on CPU in the build container:

    python tests/golden/make_golden_ends.py        ->  tests/golden/qconv2d_patch16.npz, tests/golden/qlinear4head.npz
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch
import torch.nn as nn

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as MG  # noqa: E402  (installs the reference shim, provides run_module / save)
from src.quantization.modules.qlinear import LSQ_QConv2d, LSQ_QLinear4head  # noqa: E402

torch.manual_seed(31)
conv = nn.Conv2d(3, 64, 16, 16)
q = LSQ_QConv2d(m=conv, pretrained_initialized=True)
MG.randomize_shifts(q)
MG.save("qconv2d_patch16", MG.run_module(q, torch.randn(2, 3, 224, 224)))

torch.manual_seed(32)
lin = nn.Linear(192, 1000)
q = LSQ_QLinear4head(m=lin, weight_quant_method="lsq", pretrained_initialized=True)
MG.randomize_shifts(q)
MG.save("qlinear4head", MG.run_module(q, torch.randn(8, 192)))
