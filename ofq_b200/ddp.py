"""Data-parallel gradient exchange of the QAT step (reference: train.py:727 `NativeDDP(model, device_ids=[rank])`).

The path shards by images only: every rank holds the full model, StatsQ statistics and CGA masks are
weight-only and therefore identical everywhere, and the one exchange per step is the mean of all gradients.  Instead of
DDP's bucketed hooks the gradients are gathered after backward into ONE flat fp32 buffer (a multi-tensor copy), which
is all-reduced with a single NCCL call; every `.grad` is then re-pointed at its slice of the reduced buffer, so the
optimizer reads the mean gradient from static memory. At 90.8 MB (DeiT-S) over NVLink 5 this is ~0.3 ms of a ~27 ms step,
and a single collective on static memory can be captured inside the whole-step CUDA graph. (Letting autograd accumulate
into pre-assigned views instead costs one small add kernel per parameter - ~300 launches, ~1 ms per step.)  The loss is pre-divided by the world size,
so the SUM all-reduce yields the mean gradient the reference's DDP produces.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


_ALIGN = 32          # floats


class FlatGradAllReduce:
    """`direct=True` (default): the weight-gradient GEMMs of the ofq_b200 layers write straight into the parameter's slice
    of the flat buffer (functional.GRAD_SLOTS -> take()): the slice is zeroed and handed out as a fresh view that autograd adopts
    as `.grad`, so ~94 % of the gradient bytes are never copied. Everything else (biases, norms, step sizes) is gathered by the multi-tensor copy."""

    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int, direct: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world_size = world_size
        self.direct = direct
        self._taken = set()
        dev = self.params[0].device
        # every slice starts on a 128-byte boundary (the dW GEMMs store through TMA: 16-byte aligned rows at least)
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, self.params)]
        for p, v in zip(self.params, self.views):
            p.grad = v
        self._slot = {p.data_ptr(): i for i, p in enumerate(self.params)}
        self.direct_hits = 0

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        """Start of a step: backward writes fresh gradient tensors (no accumulate kernels); reduce() gathers them."""
        for p in self.params:
            p.grad = None
        if self.direct:
            from .quantization import functional
            self._taken.clear()
            functional.GRAD_SLOTS = self

    def take(self, w: torch.Tensor):
        """A fresh zeroed view of `w`'s slice of the flat buffer for the layer's backward to accumulate dW into, once per
        step and parameter (a weight used twice in one forward gets an ordinary buffer the second time: autograd sums)."""
        i = self._slot.get(w.data_ptr())
        if i is None or i in self._taken or self.params[i].grad is not None:
            return None
        self._taken.add(i)
        self.direct_hits += 1
        p = self.params[i]
        off = self.views[i].storage_offset()
        v = self.flat[off:off + p.numel()].view_as(p)
        # zeroed HERE, right before the GEMM that accumulates into it (not by one memset of the whole buffer at the start of the
        # step: by the time a layer's backward runs those lines would have left L2 again and every reduce-add would go to HBM)
        return v.zero_()

    def scale_loss(self, loss: torch.Tensor) -> torch.Tensor:
        return loss / self.world_size if self.world_size > 1 else loss

    def reduce(self) -> None:
        grads, views = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()                                   # parameter that took no part in this step
            elif p.grad.data_ptr() != v.data_ptr():
                grads.append(p.grad)
                views.append(v)
        if grads:
            torch._foreach_copy_(views, grads)              # a handful of multi-tensor kernels
        for p, v in zip(self.params, self.views):
            p.grad = v
        if self.direct:
            from .quantization import functional
            if functional.GRAD_SLOTS is self:
                functional.GRAD_SLOTS = None
        if self.world_size > 1:
            dist.all_reduce(self.flat)


def broadcast_parameters(model: torch.nn.Module, src: int = 0) -> None:
    """DDP construction broadcast (train.py:727): also makes the lazily created, data-dependent LSQ step sizes
    identical on every rank."""
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


class BucketedGradAllReduce:
    """torch-DDP-style exchange (train.py:727): gradients are reduced in a few buckets WHILE the backward is still running.

    Parameters are grouped, in reverse registration order (the order the backward produces their gradients), into buckets
    of ~`bucket_mb`; each bucket is a slice of one flat fp32 buffer. A post-accumulate hook per parameter counts a bucket
    down; when its last gradient has arrived the bucket's gradients are gathered into the slice (one multi-tensor copy on
    the compute stream) and its NCCL all-reduce is queued on a communication stream behind an event, so the collective
    of bucket i overlaps the backward kernels of the layers below it. `reduce()` joins the communication stream (and
    handles buckets whose parameters took no part in the step). Event fork / join edges are legal inside CUDA-graph capture,
    so the whole step can still be captured as one graph. Same interface as FlatGradAllReduce."""

    def __init__(self, params_or_model, world_size: int, bucket_mb: float = 25.0):
        params = params_or_model.parameters() if isinstance(params_or_model, torch.nn.Module) else params_or_model
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world_size = world_size
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        self.comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        limit = int(bucket_mb * (1 << 20) / 4)
        self.buckets = []            # dicts: params, views, flat slice, pending count
        self._bucket_of = {}
        off = self.flat.numel()
        cur = None
        for p in reversed(self.params):
            if cur is None or cur["numel"] + p.numel() > limit:
                cur = {"params": [], "views": [], "numel": 0, "hi": off}
                self.buckets.append(cur)
            off -= p.numel()
            cur["params"].append(p)
            cur["views"].append(self.flat[off:off + p.numel()].view_as(p))
            cur["numel"] += p.numel()
            cur["lo"] = off
            self._bucket_of[id(p)] = cur
        for b in self.buckets:
            b["flat"] = self.flat[b["lo"]:b["hi"]]
            b["pending"] = len(b["params"])
            b["done"] = False
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        """Start of a step, on the stream the step runs on: the bucket gathers are queued on THAT stream (the gradients are
        produced there), whatever stream autograd happens to run a parameter's AccumulateGrad node on."""
        for p in self.params:
            p.grad = None
        for b in self.buckets:
            b["pending"], b["done"] = len(b["params"]), False
        self._main = torch.cuda.current_stream() if self.comm is not None else None

    def scale_loss(self, loss: torch.Tensor) -> torch.Tensor:
        return loss / self.world_size if self.world_size > 1 else loss

    def _on_grad(self, p) -> None:
        b = self._bucket_of[id(p)]
        b["pending"] -= 1
        if b["pending"] == 0 and not b["done"]:
            self._launch(b)

    def _gather(self, b) -> None:
        grads, views = [], []
        for p, v in zip(b["params"], b["views"]):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                grads.append(p.grad)
                views.append(v)
        if grads:
            torch._foreach_copy_(views, grads)
        for p, v in zip(b["params"], b["views"]):
            p.grad = v
        b["done"] = True

    def _launch(self, b) -> None:
        main = getattr(self, "_main", None)
        if self.comm is None or main is None:
            self._gather(b)
            if self.world_size > 1:
                dist.all_reduce(b["flat"])
            return
        with torch.cuda.stream(main):
            self._gather(b)
            if self.world_size > 1:
                self.comm.wait_event(main.record_event())
        if self.world_size > 1:
            with torch.cuda.stream(self.comm):
                dist.all_reduce(b["flat"])

    def reduce(self) -> None:
        for b in self.buckets:
            if not b["done"]:
                self._launch(b)
        if self.comm is not None and self.world_size > 1:
            torch.cuda.current_stream().wait_stream(self.comm)
