"""AdamW with the confidence-guided-annealing (CGA) weight freeze fused into the update.

Reference semantics (cga.py:953-1013 around torch.optim.AdamW, driven by timm's create_optimizer_v2):
after backward, for every masked weight a freeze mask is built from the StatsQ pre-round value
(freeze_outside_boundary_weight_idx, cga.py:450-469), the gradients of frozen weights are zeroed, the frozen
weights are stashed, AdamW steps, and the frozen weights are restored.  Here all of that is ONE kernel pass per
parameter (ofq_cga_adamw): frozen elements see a zero gradient (their moments decay exactly as AdamW-with-zero-grad)
and keep their value bit-for-bit.  With no masked parameters this is a plain fused AdamW (train.py:662, 933).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import ops


def cga_masked_parameter_names(model: torch.nn.Module, qk_reparam: bool = True, model_type: str = "deit"):
    """Weights the reference masks, by module-name suffix (cga.py:956-980)."""
    names = []
    for k, m in model.named_modules():
        if not hasattr(m, "weight") or getattr(m, "weight", None) is None or m.weight.dim() != 2:
            continue
        if qk_reparam and model_type == "swin":
            hit = k.endswith(("fc1", "fc2", ".v", "proj", "reduction"))
        elif qk_reparam:
            hit = "blocks" in k and k.endswith(("fc1", "fc2", ".v", "proj"))
        else:
            hit = "blocks" in k and k.endswith(("fc1", "fc2", "qkv", "proj"))
        if hit:
            names.append(k + ".weight")
    return names


def param_groups_weight_decay(model: torch.nn.Module, weight_decay: float, no_weight_decay=()):
    """timm 0.5.4 `add_weight_decay` (used by create_optimizer_v2, train.py:662): 1-D parameters, `.bias` and the
    names in `no_weight_decay()` get no decay."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim <= 1 or name.endswith(".bias") or name in no_weight_decay) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


class CGAAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW-compatible optimizer whose step is the fused sm_100a kernel.

    masked: iterable of parameters (2-D weights quantized by StatsQ) that take part in CGA; wq_bitw and
    boundary_range as in `cga.py --wq-bitw --boundaryRange`. Leave `masked` empty for ordinary QAT training."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, masked: Optional[Iterable] = None,
                 wq_bitw: int = 2, boundary_range: float = 0.005):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._masked = {id(p) for p in (masked or ())}
        self.wq_bitw = wq_bitw
        self.boundary_range = boundary_range
        self.last_masks = {}
        self.keep_masks = False
        self.launches = 0
        self._step_dev = None      # device-resident step counter: the whole step can be captured in a CUDA graph
        self._tables = {}          # per param-group pointer tables of the multi-tensor kernel

    # The kernels take the bias-correction step t from a device counter (so that a captured CUDA graph can be replayed);
    # it is part of the optimizer state: saved under "ofq_step" and restored (or re-seeded from the per-parameter
    # `step` entries of a torch.optim.AdamW checkpoint) by load_state_dict.
    def state_dict(self):
        sd = super().state_dict()
        sd["ofq_step"] = int(getattr(self, "_host_step", 0))
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        step = state_dict.pop("ofq_step", None)
        super().load_state_dict(state_dict)
        if step is None:
            steps = [int(st["step"]) for st in self.state.values() if "step" in st]
            step = max(steps) if steps else 0
        self._host_step = int(step)
        self._step_dev = None          # re-created on the parameters' device with the restored count at the next step
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self._step_dev is None:
            dev = next(p for g in self.param_groups for p in g["params"]).device
            self._step_dev = torch.full((1,), int(getattr(self, "_host_step", 0)), dtype=torch.int32, device=dev)
        ops.counter_increment_(self._step_dev)
        self._host_step = getattr(self, "_host_step", 0) + 1
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            plain, masked_list = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                elif torch.is_tensor(st.get("step")):
                    st["step"] = int(st["step"].item())          # a torch.optim.AdamW checkpoint keeps the step as a tensor
                st["step"] += 1
                masked = id(p) in self._masked
                if masked and "scratch" not in st:
                    st["scratch"] = (torch.empty(p.shape[0], dtype=torch.float32, device=p.device),
                                     torch.empty(2, dtype=torch.int32, device=p.device))
                dense = p.is_contiguous() and p.grad.is_contiguous()
                if dense and not masked:
                    plain.append(p)
                    continue
                if dense and not self.keep_masks:
                    masked_list.append(p)
                    continue
                # odd layouts and the mask-recording debug mode: one parameter at a time
                mask_out = None
                if masked and self.keep_masks:
                    mask_out = torch.empty(p.shape, dtype=torch.uint8, device=p.device)
                    self.last_masks[id(p)] = mask_out
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                ops.cga_adamw_(p.data, grad, st["exp_avg"], st["exp_avg_sq"], st["step"], group["lr"], b1, b2,
                               group["eps"], group["weight_decay"], bits=self.wq_bitw if masked else 0,
                               boundary_range=self.boundary_range, scratch=st.get("scratch"), mask_out=mask_out,
                               step_dev=self._step_dev)
                self.launches += 3 if masked else 1
            # Multi-tensor launches. The pointer tables are rebuilt only when a buffer moved (never under CUDA-graph
            # replay or with persistent .grad buffers); lr is a launch argument and weight decay a table field the kernel
            # turns into 1 - lr * wd, so a learning-rate schedule does not touch the tables.
            if plain:
                key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in plain) + (group["weight_decay"],)
                cache = self._tables.get((gi, "plain"))
                if cache is None or cache[0] != key:
                    entries = [(p.data, p.grad, self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"], group["weight_decay"])
                               for p in plain]
                    cache = (key,) + ops.build_adamw_table(entries, plain[0].device)
                    self._tables[(gi, "plain")] = cache
                _, table, n, blocks, numel = cache
                ops.adamw_multi_(table, n, blocks, numel, self._host_step, group["lr"], b1, b2, group["eps"],
                                 step_dev=self._step_dev)
                self.launches += 1
            if masked_list:
                key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in masked_list) + (group["weight_decay"],)
                cache = self._tables.get((gi, "masked"))
                if cache is None or cache[0] != key:
                    entries = [(p.data, p.grad, self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]) + self.state[p]["scratch"]
                               + (group["weight_decay"],) for p in masked_list]
                    cache = (key,) + ops.build_cga_table(entries, masked_list[0].device)
                    self._tables[(gi, "masked")] = cache
                _, table, n, blocks, rowblocks, numel = cache
                ops.cga_adamw_multi_(table, n, blocks, rowblocks, numel, self._host_step, group["lr"], b1, b2, group["eps"],
                                     self.wq_bitw, self.boundary_range, step_dev=self._step_dev)
                self.launches += 3
        from . import prologue
        prologue.weights_changed()         # the kernels wrote the parameters through raw pointers
        return loss
