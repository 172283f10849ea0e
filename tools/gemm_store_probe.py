import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import GEMM_I8, GEMM_F16, vec
exec(open("tools/gemm_sweep.py").read().split("for kind, name in")[0].split("dev = \"cuda\"")[1].replace("M = 25344",""))
dev="cuda"; M=25344
for N in (384, 1536, 2304):
    K=384
    A = torch.randint(-2, 2, (M, K), dtype=torch.int8, device=dev); B = torch.randint(-3, 4, (N, K), dtype=torch.int8, device=dev)
    out = torch.empty(M, N, device=dev); rs = torch.rand(198, device=dev); cs = torch.rand(N, device=dev)
    f = lambda: ops.gemm(GEMM_I8, A, (K, 0, 0, 0), B, (K, 0, 0, 0), out, (N, 0, 0), M, N, K, rs=vec(rs, 198), cs=vec(cs))
    print(f"i8 M={M} N={N} K={K}: {timeit(f):7.1f} us", flush=True)
Bt, H, Nn, C = 128, 6, 198, 384
a = torch.randn(Bt, Nn, C, device=dev).half(); b = torch.randn(Bt, Nn, C, device=dev).half()
dP = torch.empty(Bt * H, Nn, 200, device=dev)
f = lambda: ops.gemm(GEMM_F16, a, (C, 0, 64, Nn * C), b, (C, 0, 64, Nn * C), dP, (200, Nn * 200, H * Nn * 200), Nn, Nn, 64, nb1=H, nb2=Bt)
print(f"f16 dP batched M198 N198 K64: {timeit(f):7.1f} us", flush=True)
