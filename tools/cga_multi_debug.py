"""Debug: where does the multi-tensor CGA step differ from the per-tensor kernel?"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ofq_b200 import ops
from ofq_b200.cga import CGAAdamW

torch.manual_seed(18)
shapes = [(384, 384), (96, 100), (10, 7)]
ps = [torch.nn.Parameter(torch.nn.init.trunc_normal_(torch.empty(s), std=0.02).cuda()) for s in shapes]
opt = CGAAdamW([{"params": ps, "weight_decay": 0.05}], lr=1e-3, masked=ps, wq_bitw=2, boundary_range=0.05)
ref = [p.detach().clone() for p in ps]
rm = [torch.zeros_like(p) for p in ps]
rv = [torch.zeros_like(p) for p in ps]
for step in range(2):
    grads = [torch.randn_like(p) * 1e-3 for p in ps]
    for p, g in zip(ps, grads):
        p.grad = g.clone()
    opt.step()
    sd = torch.full((1,), step + 1, dtype=torch.int32, device="cuda")
    for i, (r, g) in enumerate(zip(ref, grads)):
        mask = torch.empty(r.shape, dtype=torch.uint8, device="cuda")
        before = r.clone()
        ops.cga_adamw_(r, g, rm[i], rv[i], step + 1, 1e-3, 0.9, 0.999, 1e-8, 0.05, bits=2, boundary_range=0.05, step_dev=sd, mask_out=mask)
        p = ps[i].detach()
        for name, a, b in (("p", p, r), ("m", opt.state[ps[i]]["exp_avg"], rm[i]), ("v", opt.state[ps[i]]["exp_avg_sq"], rv[i])):
            d = (a != b)
            print(step, i, name, "diff elems", int(d.sum()), "of", a.numel(), "max abs", float((a - b).abs().max()),
                  "frozen frac", float(mask.float().mean()), "diff among frozen", int((d & mask.bool()).sum()))
        st = opt.state[ps[i]]["scratch"]
        print("   kminmax", st[1].tolist(), "rowstat[0:3]", st[0][:3].tolist())
