"""Micro-benchmark of the HBM-bound kernels at the DeiT-S batch-128 shapes (CUDA events, rotating buffers > L2).
usage: python tools/kbench.py [name-substring ...]"""
import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import PER_ROW, PER_COL, FMT_F16

dev = "cuda"
want = sys.argv[1:]
M, C, H, N, B = 25344, 384, 6, 198, 128
NBUF = 4


def timeit(name, nbytes, fn, n=12):
    if want and not any(w in name for w in want):
        return
    for i in range(NBUF):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()          # one graph of n launches: no host launch overhead in the timing
    with torch.cuda.graph(graph):
        for i in range(n):
            fn(i % NBUF)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n * 1e-3
    print(f"{name:42s} {t * 1e6:8.1f} us  {nbytes / t / 1e9:7.0f} GB/s", flush=True)


def bufs(*shape, dtype=torch.float32):
    return [torch.randn(*shape, device=dev).to(dtype) if dtype.is_floating_point else
            torch.randint(-2, 2, shape, device=dev, dtype=dtype) for _ in range(NBUF)]


for cols, nseg in ((384, 1), (1536, 1), (2304, 6)):
    x = bufs(M, cols)
    dy = bufs(M, cols)
    b4 = torch.randn(cols, device=dev) * 0.1
    se = torch.rand(N * nseg, device=dev) * 0.5 + 0.2
    timeit(f"lsq_quant rows C={cols} nseg={nseg}", 5.0 * M * cols, lambda i: ops.lsq_quant(x[i], b4, se, PER_ROW, N, nseg, -2, 1))
    timeit(f"lsq_bwd rows C={cols} nseg={nseg}", 12.0 * M * cols, lambda i: ops.lsq_bwd(dy[i], x[i], b4, se, PER_ROW, N, nseg, -2, 1, 0.01))
    timeit(f"lsq_quant rows +fp16 copy C={cols} nseg={nseg}", 7.0 * M * cols, lambda i: ops.lsq_quant(x[i], b4, se, PER_ROW, N, nseg, -2, 1, fmt16=FMT_F16))
    if cols == 2304:
        u = torch.randn(cols, device=dev)
        timeit(f"lsq_quant rows +fp16 copy +rowdot C={cols} nseg={nseg}", 7.0 * M * cols, lambda i: ops.lsq_quant(x[i], b4, se, PER_ROW, N, nseg, -2, 1, fmt16=FMT_F16, dot_u=u))
    if cols == 1536:
        timeit(f"lsq_quant rows GELU +fp16 copy C={cols}", 7.0 * M * cols, lambda i: ops.lsq_quant(x[i], b4, se, PER_ROW, N, nseg, 0, 3, act=ops.ACT_GELU, fmt16=FMT_F16))
        timeit(f"lsq_bwd rows GELU C={cols}", 12.0 * M * cols, lambda i: ops.lsq_bwd(dy[i], x[i], b4, se, PER_ROW, N, nseg, 0, 3, 0.01, act=ops.ACT_GELU))
    cs = torch.rand(cols, device=dev) + 0.5
    sx = torch.rand(N, device=dev) + 0.5
    sc = ops.absmax_scale(dy[0], 1, M, cols, cols, 0, cs=cs, rs=sx, rs_period=N, product=True)
    timeit(f"absmax_scale C={cols}", 4.0 * M * cols, lambda i: ops.absmax_scale(dy[i], 1, M, cols, cols, 0, cs=cs, rs=sx, rs_period=N, product=True))
    timeit(f"grad_prep f16 rm+colsum C={cols}", 6.0 * M * cols, lambda i: ops.grad_prep(dy[i], 1, M, cols, cols, 0, cs=cs, rs=sx, rs_period=N, want_rm=True, want_colsum=True, fmt=FMT_F16, scale4=sc, rm_rowscale=True))
    q = bufs(M, cols, dtype=torch.int8)
    timeit(f"codes_to_16 C={cols}", 3.0 * M * cols, lambda i: ops.codes_to_bf16(q[i], 1, M, cols, cols, 0, False, FMT_F16))
    del x, dy, q
x = bufs(M, C)
sv = torch.rand(C, device=dev) + 0.2
b4 = torch.randn(C, device=dev) * 0.1
timeit("lsq_quant cols C=384", 5.0 * M * C, lambda i: ops.lsq_quant(x[i], b4, sv, PER_COL, 1, 1, -2, 1))
timeit("lsq_bwd cols C=384", 12.0 * M * C, lambda i: ops.lsq_bwd(x[i], x[(i + 1) % NBUF], b4, sv, PER_COL, 1, 1, -2, 1, 0.01))
g = torch.ones(C, device=dev)
timeit("layernorm_fwd", 8.0 * M * C, lambda i: ops.layernorm_fwd(x[i], g, b4, 1e-6))
y, mean, rstd = ops.layernorm_fwd(x[0], g, b4, 1e-6)
timeit("layernorm_bwd", 12.0 * M * C, lambda i: ops.layernorm_bwd(x[i], x[(i + 1) % NBUF], g, mean, rstd))
del x
nz = B * H
S = bufs(nz, N, 200)
sp = torch.full((N,), 0.02, device=dev)
timeit("softmax_quant", nz * N * N * 9.0, lambda i: ops.softmax_quant(S[i], N, H, sp, 3))
P, qp, _ = ops.softmax_quant(S[0], N, H, sp, 3)
ca = torch.rand(H, N, device=dev) + 0.5
rb = torch.rand(N, device=dev) + 0.5
sc = ops.absmax_scale(S[0], nz, N, N, 200, N * 200, v1=ca, v2=rb, mult=0.25, product=True)
timeit("absmax_scale dPq", nz * N * N * 4.0, lambda i: ops.absmax_scale(S[i], nz, N, N, 200, N * 200, v1=ca, v2=rb, mult=0.25, product=True))
timeit("softmax_quant_bwd f16 single", nz * N * N * 10.0, lambda i: ops.softmax_quant_bwd(S[i], P, N, H, sp, 3, 0.125, 0.01, ca, True, rb, fmt=FMT_F16, scale4=sc, single=True))
timeit("codes_to_16 qp", 3.0 * nz * N * 208, lambda i: ops.codes_to_bf16(qp, nz, N, 208, 208, N * 208, False, FMT_F16))
