import sys
sys.path.insert(0, ".")
import torch, torch.nn as nn
import ofq_b200.quantization as Q
from ofq_b200.host.deit import Mlp
from oracle import ofq_oracle as O
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
for shifts, sparse, scale in [(True, False, 1.0), (False, False, 1.0), (True, True, 1.0), (False, True, 1.0), (False, True, 1e-3), (False, False, 1e-3)]:
    torch.manual_seed(0)
    m = Q.QMLP(m=Mlp(64, 256), weight_bits=2, input_bits=2)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "move_" in n and shifts: p.copy_(torch.randn_like(p) * 0.05)
    x = torch.randn(2, 198, 64)
    m = m.cuda()
    with torch.no_grad(): m(x.cuda())
    xg = x.cuda().requires_grad_(True)
    y = m(xg)
    go = torch.randn(2, 198, 64) * scale
    if sparse: go[:, 2:] = 0
    y.backward(go.cuda())
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xc = x.clone().requires_grad_(True)
    yo = O.qmlp(xc, P, "", 2, 2); yo.backward(go)
    errs = sorted(((rel(p.grad.cpu(), P[n].grad), n) for n, p in m.named_parameters() if p.grad is not None and P[n].grad is not None), reverse=True)
    print(f"shifts={shifts} sparse={sparse} scale={scale}: out {rel(y.detach().cpu(), yo.detach()):.1e} dx {rel(xg.grad.cpu(), xc.grad):.1e} | " + "; ".join(f"{n} {e:.1e}" for e, n in errs[:4]))
