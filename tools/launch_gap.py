"""Cost of a kernel boundary inside a CUDA graph: n back-to-back launches of a tiny kernel, and of a streaming kernel
followed by a tiny one. usage: python tools/launch_gap.py"""
import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import PER_ROW

dev = "cuda"


def graph_time(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


amax = torch.ones(1, device=dev)
v1 = torch.rand(384, device=dev)
tiny = lambda: ops.scale_from_max(amax, v1=v1, v2=v1, product=True)
print(f"tiny kernel (scale_from_max), 200 in a graph: {graph_time(tiny, 200):6.2f} us each", flush=True)
M, C, N = 25344, 384, 198
x = torch.randn(M, C, device=dev)
b4 = torch.zeros(C, device=dev)
se = torch.rand(N, device=dev) + 0.2
big = lambda: ops.lsq_quant(x, b4, se, PER_ROW, N, 1, -2, 1)
tb = graph_time(big, 50)
print(f"lsq_quant [25344 x 384], 50 in a graph:       {tb:6.2f} us each", flush=True)


def both():
    big()
    tiny()


print(f"lsq_quant + tiny, 50 pairs in a graph:        {graph_time(both, 50):6.2f} us per pair (tiny costs the difference)", flush=True)
z = torch.empty(384, device=dev)
fill = lambda: z.zero_()
print(f"torch zero_ of 384 floats, 200 in a graph:    {graph_time(fill, 200):6.2f} us each", flush=True)
