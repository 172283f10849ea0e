"""Staged check of the attention kernels at DeiT shapes against fp64 torch math on the same codes."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from ofq_b200 import ops
from ofq_b200.ops import GEMM_I8, GEMM_BF16, vec, round_up
from conftest import rel_err
torch.manual_seed(0)
B, N, H, C = 2, 198, 2, 64
hd = C // H
M = B * N
dev = "cuda"
qq = torch.randint(-2, 2, (M, C), dtype=torch.int8, device=dev)
qk = torch.randint(-2, 2, (M, C), dtype=torch.int8, device=dev)
qv = torch.randint(-2, 2, (M, C), dtype=torch.int8, device=dev)
se_q = torch.rand(N, device=dev) + 0.5
se_k = torch.rand(N, device=dev) + 0.5
q_aft = torch.randn(C, device=dev) * 0.1
scale = hd ** -0.5
ctS = ops.codes_rowdot(qk, H, q_aft)
ref_ct = (qk.double().view(M, H, hd) * q_aft.double().view(1, H, hd)).sum(-1)
print("rowdot", rel_err(ctS, ref_ct))
cs_S = se_k * scale
ct_S = (ctS.view(B, N, H).permute(0, 2, 1) * cs_S.view(1, 1, N)).contiguous()
ldS = round_up(N, 4)
S = torch.zeros((B * H, N, ldS), device=dev)
ops.gemm(GEMM_I8, qq, (C, 0, hd, N * C), qk, (C, 0, hd, N * C), S, (ldS, N * ldS, H * N * ldS), N, N, hd, nb1=H, nb2=B,
         rs=vec(se_q, N), cs=vec(cs_S), ct=vec(ct_S, 0, N, H * N))
q4 = qq.double().view(B, N, H, hd).permute(0, 2, 1, 3)
k4 = qk.double().view(B, N, H, hd).permute(0, 2, 1, 3)
I = q4 @ k4.transpose(-1, -2)
refS = I * se_q.double().view(1, 1, N, 1) * cs_S.double().view(1, 1, 1, N) + ct_S.double().view(B, H, 1, N)
print("S plain", rel_err(S[..., :N].view(B, H, N, N), refS))
# per-tile error map
err = (S[..., :N].view(B, H, N, N).double() - refS).abs()
print(" err by (b,h):", err.amax((2, 3)).tolist())
print(" err rows>=128:", err[:, :, 128:].max().item(), " cols>=128:", err[..., 128:].max().item(), " rows<128&cols<128:", err[:, :, :128, :128].max().item())
