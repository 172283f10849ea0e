"""Per-shape timing of the GEMM engine at the DeiT-S B=128 shapes (device time, CUDA events, L2 flushed by size)."""
import sys
sys.path.insert(0, ".")
import torch
from ofq_b200 import ops
from ofq_b200.ops import GEMM_I8, GEMM_BF16, vec
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
M = 25344
print("== int8 forward linear GEMMs (out fp32)")
for (N, K) in [(384, 384), (1536, 384), (384, 1536), (2304, 384)]:
    A = torch.randint(-2, 2, (M, K), dtype=torch.int8, device=dev); B = torch.randint(-3, 4, (N, K), dtype=torch.int8, device=dev)
    out = torch.empty(M, N, device=dev); rs = torch.rand(198, device=dev); cs = torch.rand(N, device=dev); ct = torch.rand(N, device=dev)
    t = timeit(lambda: ops.gemm(GEMM_I8, A, (K, 0, 0, 0), B, (K, 0, 0, 0), out, (N, 0, 0), M, N, K, rs=vec(rs, 198), cs=vec(cs), ct=vec(ct)))
    print(f"  M={M} N={N} K={K}: {t*1e6:7.1f} us  {2*M*N*K/t/1e12:7.1f} TOP/s  bytes {(M*K+N*K+4*M*N)/t/1e9:7.0f} GB/s")
print("== bf16 backward dX GEMMs (2 planes), out fp32 [M,K]")
for (Nout, K) in [(384, 384), (1536, 384), (384, 1536), (2304, 384)]:
    A = torch.randn(2, M, Nout, device=dev).bfloat16(); B = torch.randint(-3, 4, (K, Nout), device=dev).bfloat16()
    out = torch.empty(M, K, device=dev)
    t = timeit(lambda: ops.gemm(GEMM_BF16, A, (Nout, M * Nout, 0, 0), B, (Nout, 0, 0, 0), out, (K, 0, 0), M, K, Nout, k2=2))
    print(f"  M={M} N={K} K={Nout}x2: {t*1e6:7.1f} us  {2*M*K*Nout*2/t/1e12:7.1f} TFLOP/s  bytes {(2*2*M*Nout+2*K*Nout+4*M*K)/t/1e9:7.0f} GB/s")
print("== bf16 backward dW GEMMs (2 planes, split-K), out fp32 [Nout,K]")
for (Nout, K) in [(384, 384), (1536, 384), (384, 1536), (2304, 384)]:
    A = torch.randn(2, Nout, M, device=dev).bfloat16(); B = torch.randint(-3, 4, (K, M), device=dev).bfloat16()
    out = torch.zeros(Nout, K, device=dev)
    tiles = ((Nout + 127) // 128) * ((K + 127) // 128)
    splits = max(1, min((2 * 148 + tiles - 1) // tiles, (2 * ((M + 63) // 64)) // 4, 64))
    t = timeit(lambda: ops.gemm(GEMM_BF16, A, (M, Nout * M, 0, 0), B, (M, 0, 0, 0), out, (K, 0, 0), Nout, K, M, k2=2, splits=splits, accumulate=True))
    print(f"  M={Nout} N={K} K={M}x2 splits={splits}: {t*1e6:7.1f} us  {2*M*K*Nout*2/t/1e12:7.1f} TFLOP/s  bytes {(2*2*M*Nout+2*K*M)/t/1e9:7.0f} GB/s")
print("== attention GEMMs B=128 H=6 N=198")
Bt, H, N, C = 128, 6, 198, 384
qx = torch.randint(-2, 2, (Bt, N, C), dtype=torch.int8, device=dev); qk = torch.randint(-2, 2, (Bt, N, H, C), dtype=torch.int8, device=dev)
S = torch.empty(Bt * H, N, 200, device=dev)
t = timeit(lambda: ops.gemm(GEMM_I8, qx, (C, 0, 0, N * C), qk, (H * C, 0, C, N * H * C), S, (200, N * 200, H * N * 200), N, N, C, nb1=H, nb2=Bt))
print(f"  scores i8 K=384: {t*1e6:7.1f} us {2*Bt*H*N*N*C/t/1e12:7.1f} TOP/s  out {Bt*H*N*200*4/t/1e9:7.0f} GB/s")
qp = torch.randint(0, 4, (Bt * H, N, 208), dtype=torch.int8, device=dev); qvT = torch.randint(-2, 2, (Bt, C, 208), dtype=torch.int8, device=dev)
o = torch.empty(Bt, N, C, device=dev)
t = timeit(lambda: ops.gemm(GEMM_I8, qp, (208, 0, N * 208, H * N * 208), qvT, (208, 0, 64 * 208, C * 208), o, (C, 64, N * C), N, 64, N, nb1=H, nb2=Bt))
print(f"  P.V i8: {t*1e6:7.1f} us {2*Bt*H*N*N*64/t/1e12:7.1f} TOP/s  in {(Bt*H*N*208+Bt*C*208)/t/1e9:7.0f} GB/s")
dSa = torch.randn(Bt, 2, H, N, 200, device=dev).bfloat16(); qkT = torch.randint(-2, 2, (Bt, H * C, 200), device=dev).bfloat16()
dx = torch.zeros(Bt, N, C, device=dev); slab = N * 200
t = timeit(lambda: ops.gemm(GEMM_BF16, dSa, (200, slab, 2 * H * slab, 0), qkT, (200, C * 200, H * C * 200, 0), dx, (C, N * C, 0), N, C, N, k2=2 * H, nb1=Bt, accumulate=True, b_k2mod=H))
print(f"  dx_hat from scores bf16 (k2=12): {t*1e6:7.1f} us {2*Bt*H*N*N*C*2/t/1e12:7.1f} TFLOP/s")
qxT = torch.randint(-2, 2, (Bt, C, 200), device=dev).bfloat16(); dk = torch.empty(Bt * N, H * C, device=dev)
t = timeit(lambda: ops.gemm(GEMM_BF16, dSa, (200, H * slab, slab, 2 * H * slab), qxT, (200, 0, 0, C * 200), dk, (H * C, C, N * H * C), N, C, N, k2=2, nb1=H, nb2=Bt))
print(f"  dk_hat bf16: {t*1e6:7.1f} us {2*Bt*H*N*N*C*2/t/1e12:7.1f} TFLOP/s  out {Bt*N*H*C*4/t/1e9:7.0f} GB/s")
