"""Per-phase clock totals of the GEMM engine (library built with OFQ_NVCC_FLAGS=-DOFQ_GEMM_TRACE; single-CTA kernel:
OFQ_GEMM_PAIR=0). usage: OFQ_GEMM_PAIR=0 python tools/gemm_trace.py [N ...]"""
import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import GEMM_I8, GEMM_F16, vec

dev = "cuda"
M, K = 25344, 384
for N in [int(a) for a in sys.argv[1:]] or [2304, 384]:
    A = torch.randint(-2, 2, (M, K), dtype=torch.int8, device=dev)
    B = torch.randint(-3, 4, (N, K), dtype=torch.int8, device=dev)
    out = torch.empty(M, N, device=dev)
    rs = torch.rand(198, device=dev)
    cs = torch.rand(N, device=dev)
    ct = torch.rand(N, device=dev)
    for i in range(3):
        print(f"--- i8 M={M} N={N} K={K} launch {i}", flush=True)
        ops.gemm(GEMM_I8, A, (K, 0, 0, 0), B, (K, 0, 0, 0), out, (N, 0, 0), M, N, K, rs=vec(rs, 198), cs=vec(cs), ct=vec(ct))
        torch.cuda.synchronize()

# weight-gradient shape (both operands MN-major, auto split-K): dW[Mo, No] = A16[T, Mo]^T @ Q16[T, No]
T = 25344
for Mo, No in ((384, 1536), (2304, 384)):
    a16 = torch.randn(T, Mo, device=dev).half()
    q16 = torch.randint(-2, 2, (T, No), device=dev).half()
    dW = torch.zeros(Mo, No, device=dev)
    for i in range(2):
        print(f"--- f16 dW M={Mo} N={No} K={T} launch {i}", flush=True)
        ops.gemm(GEMM_F16, a16, (Mo, 0, 0, 0), q16, (No, 0, 0, 0), dW, (No, 0, 0), Mo, No, T, a_mn=True, b_mn=True,
                 splits=0, accumulate=True)
        torch.cuda.synchronize()
