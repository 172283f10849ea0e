import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, torch.nn as nn
import ofq_b200.quantization as Q
from ofq_b200.host.deit import Attention
from oracle import ofq_oracle as O
from conftest import rel_err
torch.manual_seed(0)
for cls, fn in ((Q.QAttention_qkreparam, O.qattention_qkr), (Q.QAttention, O.qattention)):
  for (B, N, C, H) in [(2, 10, 32, 2), (2, 198, 64, 2), (2, 128, 64, 2), (2, 64, 64, 2), (3, 72, 64, 1), (2, 198, 384, 6)]:
    m = cls(Attention(C, H, qkv_bias=True), weight_bits=2, input_bits=2)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "move_" in n: p.copy_(torch.randn_like(p) * 0.05)
    x = torch.randn(B, N, C)
    m = m.cuda()
    with torch.no_grad():
        m(x.cuda())
    xg = x.cuda().requires_grad_(True)
    y, _ = m(xg)
    go = torch.randn(B, N, C)
    y.backward(go.cuda())
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xc = x.clone().requires_grad_(True)
    yo = fn(xc, P, "", H, 2, 2)
    yo.backward(go)
    print(cls.__name__, B, N, C, H, "out", f"{rel_err(y.detach().cpu(), yo.detach()):.1e}", "dx", f"{rel_err(xg.grad.cpu(), xc.grad):.1e}")
    errs = []
    for n, p in m.named_parameters():
        if P[n].grad is None: continue
        errs.append((rel_err(p.grad.cpu(), P[n].grad), n, P[n].grad.norm().item()))
    errs.sort(reverse=True)
    print("    " + "; ".join(f"{n} {e:.1e}" for e, n, _ in errs[:6]))
