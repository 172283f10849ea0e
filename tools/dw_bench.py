"""Weight-gradient GEMM shapes of DeiT-S (both operands MN-major, auto split-K), CUDA-graph timed with rotating operands.
usage: [OFQ_GEMM_PAIR=0|2] python tools/dw_bench.py"""
import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import GEMM_F16

dev = "cuda"
T = 25344
NBUF = 3


def timeit(fn, n=12):
    for i in range(NBUF):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i % NBUF)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for Mo, No in ((384, 384), (384, 1536), (1536, 384), (2304, 384)):
    a16 = [torch.randn(T, Mo, device=dev).half() for _ in range(NBUF)]
    q16 = [torch.randint(-2, 2, (T, No), device=dev).half() for _ in range(NBUF)]
    dW = torch.zeros(Mo, No, device=dev)
    f = lambda i: ops.gemm(GEMM_F16, a16[i], (Mo, 0, 0, 0), q16[i], (No, 0, 0, 0), dW, (No, 0, 0), Mo, No, T, a_mn=True,
                           b_mn=True, splits=0, accumulate=True)
    t = timeit(f)
    print(f"dW M={Mo:5d} N={No:5d} K={T}: {t:7.1f} us  {2.0 * Mo * No * T / t / 1e6:7.1f} TFLOP/s", flush=True)
