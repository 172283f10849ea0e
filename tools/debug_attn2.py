import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, torch.nn as nn
import ofq_b200.quantization as Q
from ofq_b200.host.deit import Attention
from oracle import ofq_oracle as O
from conftest import rel_err
torch.manual_seed(0)
for (B, N, C, H) in [(2, 10, 32, 2), (2, 198, 64, 2), (2, 130, 64, 2), (2, 128, 64, 2), (2, 198, 32, 2), (2, 64, 64, 2)]:
    m = Q.QAttention(Attention(C, H, qkv_bias=True), weight_bits=2, input_bits=2)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "move_" in n: p.copy_(torch.randn_like(p) * 0.05)
    x = torch.randn(B, N, C)
    m = m.cuda()
    xg = x.cuda().requires_grad_(True)
    y, _ = m(xg)          # lazily inits scales from data
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xc = x.clone().requires_grad_(True)
    yo = O.qattention(xc, P, "", H, 2, 2)
    print(B, N, C, H, "out rel", rel_err(y.detach().cpu(), yo.detach()))
    qkv = m.qkv(xg)
    qkvo = O.qlinear(xc, P, "qkv.", 2, 2)
    print("   qkv rel", rel_err(qkv.detach().cpu(), qkvo.detach()))
    core = m._core(qkv)
    # oracle core
    import torch.nn.functional as F
    hd = C // H
    t = (qkvo + P["move_qkv_b4.bias"]).reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    q = O.lsq_rows(q, P["quan_a_q_fn.s"], 2, False); k = O.lsq_rows(k, P["quan_a_k_fn.s"], 2, False)
    v = O.lsq_cols(v.permute(0, 2, 1, 3).reshape(B, N, C), P["quan_a_v_fn.s"], 2, False)
    q = (q.permute(0, 2, 1, 3).reshape(B, N, C) + P["move_q_aft.bias"]).reshape(B, N, H, hd).permute(0, 2, 1, 3)
    k = (k.permute(0, 2, 1, 3).reshape(B, N, C) + P["move_k_aft.bias"]).reshape(B, N, H, hd).permute(0, 2, 1, 3)
    v = (v + P["move_v_aft.bias"]).reshape(B, N, H, hd).permute(0, 2, 1, 3)
    attn = (q @ k.transpose(-2, -1)) * hd ** -0.5
    prob = O.lsq_rows(F.softmax(attn, -1), P["quan_a_softmax_fn.s"], 2, True)
    co = (prob @ v).transpose(1, 2).reshape(B, N, C)
    print("   core rel", rel_err(core.detach().cpu(), co.detach()))
