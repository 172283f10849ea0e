"""The knowledge-distillation teacher of the training recipe (reference train.py:428-442 create_teacher_model, 896-910).

Every script under the reference's train_scripts/ trains with `--use-kd --kd_hard_and_soft 1`: an unquantized, pretrained
DeiT / Swin of the same family produces soft targets for each batch. The reference calls `teacher(input)` inside the
autograd graph of the step (with the teacher left in training mode, so a distilled DeiT returns its (class, distillation)
logit pair and the loss takes the class logits) and back-propagates into it for nothing; here the teacher is frozen, runs
under `torch.no_grad()` and returns the same logits. `dtype=torch.bfloat16` runs it under autocast on the bf16 tensor cores
(library GEMMs / SDPA: the teacher is outside the quantized hot path) - faster, but the soft targets then carry bf16
rounding, so the parity default is fp32.
"""
from __future__ import annotations

from typing import Optional

import torch


class Teacher:
    def __init__(self, model: torch.nn.Module, dtype: Optional[torch.dtype] = None):
        self.model = model
        self.dtype = dtype
        for p in model.parameters():
            p.requires_grad_(False)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor):
        """Returns what `soft_target, _ = teacher(input)` binds in train.py:900/906: the model's first output."""
        if self.dtype is not None and self.dtype != torch.float32:
            with torch.autocast("cuda", dtype=self.dtype):
                out = self.model(x)
        else:
            out = self.model(x)
        # the reference's DeiT / Swin hosts return (logits, attention info); a bare (class, distillation) pair is kept whole
        if isinstance(out, tuple) and len(out) == 2 and not torch.is_tensor(out[1]):
            out = out[0]
        if isinstance(out, tuple):
            return tuple(o.float() for o in out)
        return out.float()
