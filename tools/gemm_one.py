"""Launch one GEMM shape a few times (for ncu)."""
import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import GEMM_I8, GEMM_BF16, vec
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
if which == "scores_i8":
    Bt, H, N, C = 128, 6, 198, 384
    qx = torch.randint(-2, 2, (Bt, N, C), dtype=torch.int8, device=dev); qk = torch.randint(-2, 2, (Bt, N, H, C), dtype=torch.int8, device=dev)
    S = torch.empty(Bt * H, N, 200, device=dev)
    f = lambda: ops.gemm(GEMM_I8, qx, (C, 0, 0, N * C), qk, (H * C, 0, C, N * H * C), S, (200, N * 200, H * N * 200), N, N, C, nb1=H, nb2=Bt)
for _ in range(5): f()
torch.cuda.synchronize()
