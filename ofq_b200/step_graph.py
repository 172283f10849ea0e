"""Whole-step CUDA graph capture as a product helper.

The QAT step of the hot path is ~1 000 short kernels; launched from Python one by one the host, not the GPU, sets the pace.
`CapturedStep` records ONE call of a step function (forward + loss + backward [+ gradient all-reduce] + optimizer) on
static input buffers into a CUDA graph and replays it:

    step = CapturedStep(train_step, (images, labels))     # eager warm-up on a side stream, then capture
    loss = step(images, labels)                            # copies the new batch into the static buffers, replays

Requirements on `fn` (all met by the ofq_b200 modules + CGAAdamW + ddp.*GradAllReduce): no host synchronisation, no
data-dependent Python control flow, optimizer state created before capture (the warm-up calls do that), step counter on the
device (CGAAdamW), gradients either freshly produced each step (`zero_grad(set_to_none=True)`) or static views.
The reference's train.py drives an eager loop; wrapping its `model(input) ... optimizer.step()` body in this helper is the one
change that lets the drop-in modules run at kernel speed (bench.py reports both: `--graph on|off`).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class CapturedStep:
    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 2, pre_capture: Callable = None):
        self.fn = fn
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if pre_capture is not None:
            pre_capture()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_output = fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_output

    def replay(self):
        self.graph.replay()
        return self.static_output
