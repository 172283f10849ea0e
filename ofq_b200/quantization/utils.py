"""KD losses of the training recipe: mirror of the reference's `src/quantization/utils.py:30-77` (same class names, call
signatures and tuple handling), computed with their gradients by ONE kernel (`ofq_kd_loss`).

    train.py:754-758   train_loss_fn = KLLossSoft()              (--kd_hard_and_soft 0)
                       train_loss_fn = KDLossSoftandHard()       (--kd_hard_and_soft 1, every script under train_scripts/)
    train.py:896-910   loss = loss_fn(student_logit, soft_target) / loss_fn(student_logit, target, soft_target)

The teacher's logits are a constant of the step (the reference back-propagates into the teacher and discards the result;
`ofq_b200.kd.Teacher` runs it under no_grad). Hard targets are class indices; probability targets (mixup) fall back to the
torch composition, as do reductions other than "mean".
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


def _first(x):
    return x[0] if isinstance(x, tuple) else x


class _KDLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_hard, z_soft, teacher, target, T: float):
        loss, dzh, dzs = ops.kd_loss(z_hard, z_soft, teacher, target, T)
        ctx.save_for_backward(dzh, dzs)
        ctx.same = z_hard is not None and z_soft is not None and z_hard.data_ptr() == z_soft.data_ptr()
        return loss

    @staticmethod
    def backward(ctx, g):
        dzh, dzs = ctx.saved_tensors
        gh = None if dzh is None else dzh * g
        gs = None if (dzs is None or ctx.same) else dzs * g
        return gh, gs, None, None, None


def _fusable(*ts):
    return all(t is None or (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2) for t in ts)


class KLLossSoft(torch.nn.modules.loss._Loss):
    """-sum softmax(target / T) * log_softmax(output / T), mean over the batch (utils.py:44-58)."""

    def forward(self, output, target, T=1.0):
        output, target = _first(output), _first(target)
        if self.reduction == "mean" and _fusable(output, target):
            return _KDLossFn.apply(None, output.contiguous(), target.detach().contiguous(), None, float(T))
        output, target = output / T, target / T
        loss = -torch.sum(F.softmax(target, dim=1) * F.log_softmax(output, dim=1), dim=1)
        return loss.mean() if self.reduction == "mean" else (loss.sum() if self.reduction == "sum" else loss)


Multi_KLLossSoft = KLLossSoft          # utils.py:30-42: the same computation under another name


class KDLossSoftandHard(nn.Module):
    """Hard cross entropy on the class head + KLLossSoft of the distillation head against the teacher (utils.py:60-77)."""

    def __init__(self) -> None:
        super().__init__()
        self.KLSoft = KLLossSoft()
        self.Hard = nn.CrossEntropyLoss()

    def forward(self, output, hard_target, soft_target):
        cls_output, dist_output = (output[0], output[1]) if isinstance(output, tuple) else (output, output)
        soft_target = _first(soft_target)
        if hard_target.dtype == torch.int64 and hard_target.dim() == 1 and _fusable(cls_output, dist_output, soft_target):
            cls_c = cls_output.contiguous()
            dist_c = cls_c if dist_output is cls_output else dist_output.contiguous()
            return _KDLossFn.apply(cls_c, dist_c, soft_target.detach().contiguous(), hard_target, 1.0)
        return self.KLSoft(dist_output, soft_target) + self.Hard(cls_output, hard_target)
