"""Stand-alone timing of ofq_gemm_lsq (qkx GEMM with the quantizer as its epilogue) against ofq_gemm + ofq_lsq_quant_ex at the
DeiT-S bench shape (M = 128 * 198, N = 6 * 384, K = 384). CUDA events, L2 flushed between repetitions.
    python tools/gemm_lsq_bench.py [B]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ofq_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N, C, H, bits = 198, 384, 6, 2
M = B * N
lo, hi = -2, 1
dev = "cuda"
torch.manual_seed(0)
qx = torch.randint(lo, hi + 1, (M, C), dtype=torch.int8, device=dev)
wc = (torch.randint(lo, hi + 1, (H * C, C), device=dev) * 2 + 1).to(torch.int8)
se_x = torch.rand(N, device=dev) * 0.2 + 0.1
cs = torch.rand(H * C, device=dev) * 0.02 + 0.01
ct = torch.randn(H * C, device=dev) * 0.05
b4 = torch.randn(H * C, device=dev) * 0.05
u = torch.randn(H * C, device=dev) * 0.1
s2 = ops.lsq_effective_scale(torch.rand(N * H, device=dev) * 0.5 + 0.2, 0.01, recip=True)
y = torch.empty((M, H * C), dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
vec = ops.vec


def timeit(fn, reps=10):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def unfused():
    ops.gemm(ops.GEMM_I8, qx, (C, 0, 0, 0), wc, (C, 0, 0, 0), y, (H * C, 0, 0), M, H * C, C, rs=vec(se_x, N), cs=vec(cs), ct=vec(ct))
    return ops.lsq_quant(y, b4, s2[0], ops.PER_ROW, N, H, lo, hi, fmt16=ops.FMT_F16, dot_u=u)


def gemm_only():
    ops.gemm(ops.GEMM_I8, qx, (C, 0, 0, 0), wc, (C, 0, 0, 0), y, (H * C, 0, 0), M, H * C, C, rs=vec(se_x, N), cs=vec(cs), ct=vec(ct))


def fused(fmt16=ops.FMT_F16, res=True, dot=True):
    return ops.gemm_lsq(qx, wc, M, H * C, C, b4, s2, N, H, lo, hi, rs=vec(se_x, N), cs=vec(cs), ct=vec(ct), fmt16=fmt16, want_res=res,
                        dot_u=u if dot else None)


print(f"M={M} N={H * C} K={C}")
print(f"ofq_gemm (fp32 out)                  {timeit(gemm_only):8.1f} us")
print(f"ofq_gemm + ofq_lsq_quant_ex          {timeit(unfused):8.1f} us")
print(f"ofq_gemm_lsq codes+fp16+res+dot      {timeit(fused):8.1f} us")
print(f"ofq_gemm_lsq codes+fp16+dot          {timeit(lambda: fused(res=False)):8.1f} us")
print(f"ofq_gemm_lsq codes+dot               {timeit(lambda: fused(None, False)):8.1f} us")
print(f"ofq_gemm_lsq codes                   {timeit(lambda: fused(None, False, False)):8.1f} us")
