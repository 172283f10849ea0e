"""Per-k-block cost of the GEMM engine: time vs K at fixed M, N (CUDA graph of n launches, rotating operands > L2 off:
operands stay L2-resident on purpose when --hot). usage: [OFQ_GEMM_PAIR=0] python tools/gemm_sweep.py"""
import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import GEMM_I8, GEMM_F16, vec

dev = "cuda"
M = 25344


def timeit(fn, n=10):
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


for kind, name in ((GEMM_I8, "i8 "), (GEMM_F16, "f16"), (GEMM_F16, "f16 bT")):
    for N in (192, 384, 1536):
        row = []
        for K in (256, 1024, 2048, 4096):
            if kind == GEMM_I8:
                A = torch.randint(-2, 2, (M, K), dtype=torch.int8, device=dev)
                B = torch.randint(-3, 4, (N, K), dtype=torch.int8, device=dev)
            else:
                A = torch.randn(M, K, device=dev).half()
                B = torch.randn(K, N, device=dev).half() if name.endswith("bT") else torch.randn(N, K, device=dev).half()
            out = torch.empty(M, N, device=dev)
            rs = torch.rand(198, device=dev)
            cs = torch.rand(N, device=dev)
            bmn = name.endswith("bT")
            f = lambda: ops.gemm(kind, A, (K, 0, 0, 0), B, ((N if bmn else K), 0, 0, 0), out, (N, 0, 0), M, N, K, b_mn=bmn,
                                 rs=vec(rs, 198), cs=vec(cs))
            t = timeit(f)
            row.append((K, t))
            del A, B, out
        kb = 128 if kind == GEMM_I8 else 64
        (k0, t0), (k1, t1) = row[1], row[3]
        slope_us_per_kblock = (t1 - t0) / ((k1 - k0) / kb)
        print(f"{name:7s} N={N:5d} " + " ".join(f"K={k}:{t:7.1f}us" for k, t in row) +
              f"  | {slope_us_per_kblock * 1e3:7.1f} ns per k-block column, TFLOP/s at K=4096: {2.0 * M * N * 4096 / row[3][1] / 1e6:7.1f}", flush=True)
