"""Stand-alone timing of the fused QKR attention forward (ofq_qkr_attn_fwd) at the DeiT-S bench shape against the
three-kernel path, with the optional outputs switched on / off. CUDA events, L2 flushed between repetitions.
    python tools/attn_bench.py [B] [H]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from ofq_b200 import ops  # noqa: E402
from ofq_b200.quantization import functional as Fn  # noqa: E402
from test_gpu_attn_fused import _unfused  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H = int(sys.argv[2]) if len(sys.argv) > 2 else 6
N, C, bits = 198, 64 * H, 2
dev = "cuda"
torch.manual_seed(0)
lo, hi, qhi = -2, 1, 3
qx = torch.randint(lo, hi + 1, (B * N, C), dtype=torch.int8, device=dev)
qk = torch.randint(lo, hi + 1, (B * N, H * C), dtype=torch.int8, device=dev)
qv = torch.randint(lo, hi + 1, (B * N, C), dtype=torch.int8, device=dev)
se_x = torch.rand(N, device=dev) * 0.5 + 0.5
se_k = (torch.rand(N * H, device=dev) * 0.5 + 0.5) * (2.0 / C ** 0.5)
ctS = torch.randn(B * N, H, device=dev)
se_p = torch.rand(N, device=dev) * 0.01 + 0.008
se_v = torch.rand(C, device=dev) * 0.1 + 0.05
v_aft = torch.randn(C, device=dev) * 0.02
scale = 0.125
qvT = ops.codes_transpose(qv, B, N, C, C, N * C)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


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


for name, kw in (("codes + out only (eval)", {}), ("+ qp16", dict(fmt16=ops.FMT_F16)), ("+ P fp32", dict(save_p=True)),
                 ("+ P + qp16 (training, round-2 backward)", dict(save_p=True, fmt16=ops.FMT_F16))):
    t = timeit(lambda: ops.qkr_attn_fwd(qx, qk, qvT, B, N, H, C, se_x, se_k, ctS, scale, se_p, qhi, se_v, v_aft, **kw))
    print(f"fused   {t:8.1f} us  {name}")
t = timeit(lambda: _unfused(ops, Fn, qx, qk, qv, B, N, H, C, se_x, se_k, ctS, scale, se_p, qhi, se_v, v_aft, ops.FMT_F16))
print(f"unfused {t:8.1f} us  score GEMM + softmax_quant (P, codes, qp16) + codes_transpose + P.V GEMM")

# ---- backward: fused kernel against dP GEMM + softmax_quant_bwd
from ofq_b200.ops import FMT_F16  # noqa: E402
sp2 = torch.stack((se_p, 1.0 / se_p)).contiguous()
sv2 = torch.stack((se_v, 1.0 / se_v)).contiguous()
g_p = 1.0 / ((qhi * B * H * N) ** 0.5)
dO = torch.randn(B, N, C, device=dev) * 1e-3
out, qp, P, qp16, _, rowstat = ops.qkr_attn_fwd(qx, qk, qvT, B, N, H, C, se_x, se_k, ctS, scale, se_p, qhi, se_v, v_aft, save_p=True,
                                                fmt16=FMT_F16, want_rowstat=True)
ldq, ldS = qp.shape[-1], P.shape[-1]
(a16, rowdot, sc_in, qv16), _ = Fn._pv_backward_f16(dO, qp, ldq, qv, sp2, sv2, v_aft, B, N, H, C, ldS, None, qp16, None, skip_dp=True)
se_k_hn = se_k.view(N, H).t().contiguous()
t = timeit(lambda: ops.qkr_attn_bwd(qx, qk, a16, qv16, FMT_F16, B, N, H, C, se_x, se_k, ctS, scale, sp2, qhi, rowstat, rowdot, sc_in,
                                    se_v, v_aft, 2, g_p))
print(f"fused   {t:8.1f} us  backward: S recompute + dP + softmax / quantizer backward -> dS16, colsum, ds")


def unfused_bwd():
    amax = torch.zeros(1, device=dev)
    dPq = torch.empty((B * H, N, ldS), dtype=torch.float32, device=dev)
    ops.gemm(ops.GEMM_F16, a16, (C, 0, 64, N * C), qv16, (C, 0, 64, N * C), dPq, (ldS, N * ldS, H * N * ldS), N, N, 64, nb1=H, nb2=B,
             rs=ops.vec(sp2[1], N), cs=ops.vec(sc_in[1:2], 1), rt=ops.vec(rowdot, 0, N, H * N), amax=amax)
    sc = ops.scale_from_max(amax, v1=se_k_hn, v2=se_x, mult=2.0 * scale, product=True)
    ops.softmax_quant_bwd(dPq, P, N, H, se_p, qhi, scale, g_p, se_k_hn, True, se_x, fmt=FMT_F16, scale4=sc, single=True)


t = timeit(unfused_bwd)
print(f"unfused {t:8.1f} us  backward: dP GEMM (fp32 out) + scale + softmax_quant_bwd on stored P")
