"""Launch one GEMM shape a few times (for ncu)."""
import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import GEMM_I8, GEMM_BF16, GEMM_F16, vec
which = sys.argv[1] if len(sys.argv) > 1 else "fc1_i8"
dev = "cuda"; M = 25344
if which == "fc1_i8":
    N, K = 1536, 384
    A = torch.randint(-2, 2, (M, K), dtype=torch.int8, device=dev); B = torch.randint(-3, 4, (N, K), dtype=torch.int8, device=dev)
    out = torch.empty(M, N, device=dev); rs = torch.rand(198, device=dev); cs = torch.rand(N, device=dev); ct = torch.rand(N, device=dev)
    f = lambda: ops.gemm(GEMM_I8, A, (K, 0, 0, 0), B, (K, 0, 0, 0), out, (N, 0, 0), M, N, K, rs=vec(rs, 198), cs=vec(cs), ct=vec(ct))
elif which == "fc2_dx_bf16":
    Nout, K = 384, 1536   # dX of fc2: [M,1536] = dY[M,384] (2 planes) x codes^T
    A = torch.randn(2, M, Nout, device=dev).bfloat16(); B = torch.randint(-3, 4, (K, Nout), device=dev).bfloat16()
    out = torch.empty(M, K, device=dev)
    f = lambda: ops.gemm(GEMM_BF16, A, (Nout, M * Nout, 0, 0), B, (Nout, 0, 0, 0), out, (K, 0, 0), M, K, Nout, k2=2)
elif which == "fc1_dx_bf16":
    Nout, K = 1536, 384
    A = torch.randn(2, M, Nout, device=dev).bfloat16(); B = torch.randint(-3, 4, (K, Nout), device=dev).bfloat16()
    out = torch.empty(M, K, device=dev)
    f = lambda: ops.gemm(GEMM_BF16, A, (Nout, M * Nout, 0, 0), B, (Nout, 0, 0, 0), out, (K, 0, 0), M, K, Nout, k2=2)
elif which == "fc1_dx_f16":     # dX of fc1 in the default fp16 mode: [M,384] = A16[M,1536] x codes[1536,384] (B MN-major)
    Nout, K = 1536, 384
    A = (torch.randn(M, Nout, device=dev) * 100).half(); B = torch.randint(-3, 4, (Nout, K), device=dev).half()
    out = torch.empty(M, K, device=dev); rs = torch.rand(198, device=dev); sc = torch.rand(1, device=dev)
    f = lambda: ops.gemm(GEMM_F16, A, (Nout, 0, 0, 0), B, (K, 0, 0, 0), out, (K, 0, 0), M, K, Nout, b_mn=True, rs=vec(rs, 198), cs=vec(sc, 1))
elif which == "fc1_dw_f16":     # dW of fc1: [1536,384] = A16^T[1536,M] x qx16[M,384] (both MN-major), split-K 9
    Nout, K = 1536, 384
    A = (torch.randn(M, Nout, device=dev) * 100).half(); B = torch.randint(-2, 2, (M, K), device=dev).half()
    out = torch.zeros(Nout, K, device=dev); cs = torch.rand(1, device=dev); rs = torch.rand(Nout, device=dev)
    f = lambda: ops.gemm(GEMM_F16, A, (Nout, 0, 0, 0), B, (K, 0, 0, 0), out, (K, 0, 0), Nout, K, M, a_mn=True, b_mn=True, splits=9,
                         accumulate=True, rs=vec(rs), cs=vec(cs, 1))
elif which == "qkx_i8":         # qkx = x_hat W_qk^T: [M, 2304] int8 x int8, K = 384 (output-write bound)
    N, K = 2304, 384
    A = torch.randint(-2, 2, (M, K), dtype=torch.int8, device=dev); B = torch.randint(-3, 4, (N, K), dtype=torch.int8, device=dev)
    out = torch.empty(M, N, device=dev); rs = torch.rand(198, device=dev); cs = torch.rand(N, device=dev); ct = torch.rand(N, device=dev)
    f = lambda: ops.gemm(GEMM_I8, A, (K, 0, 0, 0), B, (K, 0, 0, 0), out, (N, 0, 0), M, N, K, rs=vec(rs, 198), cs=vec(cs), ct=vec(ct))
elif which == "fc2_i8":         # fc2 forward: [M,384] = codes[M,1536] x codes[384,1536]^T (tensor bound)
    N, K = 384, 1536
    A = torch.randint(0, 4, (M, K), dtype=torch.int8, device=dev); B = torch.randint(-3, 4, (N, K), dtype=torch.int8, device=dev)
    out = torch.empty(M, N, device=dev); rs = torch.rand(198, device=dev); cs = torch.rand(N, device=dev); ct = torch.rand(N, device=dev)
    f = lambda: ops.gemm(GEMM_I8, A, (K, 0, 0, 0), B, (K, 0, 0, 0), out, (N, 0, 0), M, N, K, rs=vec(rs, 198), cs=vec(cs), ct=vec(ct))
if which == "scores_i8":
    Bt, H, N, C = 128, 6, 198, 384
    qx = torch.randint(-2, 2, (Bt, N, C), dtype=torch.int8, device=dev); qk = torch.randint(-2, 2, (Bt, N, H, C), dtype=torch.int8, device=dev)
    S = torch.empty(Bt * H, N, 200, device=dev)
    f = lambda: ops.gemm(GEMM_I8, qx, (C, 0, 0, N * C), qk, (H * C, 0, C, N * H * C), S, (200, N * 200, H * N * 200), N, N, C, nb1=H, nb2=Bt)
for _ in range(5): f()
torch.cuda.synchronize()
