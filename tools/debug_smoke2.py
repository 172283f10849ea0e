import sys
sys.path.insert(0, ".")
import torch, torch.nn.functional as F
import ofq_b200.quantization as Q
from ofq_b200.host.deit import DistilledVisionTransformer
from oracle import ofq_oracle as O
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
torch.manual_seed(0)
depth, dim, heads = 2, 64, 2
model = DistilledVisionTransformer(embed_dim=dim, depth=depth, num_heads=heads, num_classes=10)
names = Q.deit_qmodule_names(depth)
model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(names, 2, 2), pretrained_initialized=True, qk_reparam=True, qk_reparam_type=1).cuda()
img = torch.randn(2, 3, 224, 224); labels = torch.tensor([1, 5])
model.eval()
with torch.no_grad(): model(img.cuda())
model.train()
cap = {}
mlp = model.blocks[1].mlp
def fwd_hook(mod, inp, out):
    cap["h"] = inp[0].detach().clone()
    out.register_hook(lambda g: cap.__setitem__("go", g.detach().clone()))
mlp.register_forward_hook(fwd_hook)
(cls, dist), _ = model(img.cuda())
loss = F.cross_entropy(cls, labels.cuda()) + F.cross_entropy(dist, labels.cuda()); loss.backward()
in_model = {n: p.grad.detach().cpu().clone() for n, p in mlp.named_parameters() if p.grad is not None}
# oracle on captured input / upstream gradient
P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in mlp.state_dict().items()}
import math
hc = cap["h"].cpu().clone().requires_grad_(True)
yo = O.qmlp(hc, P, "", 2, 2); yo.backward(cap["go"].cpu())
print("upstream grad: nonzero rows", (cap["go"].abs().sum(-1) > 0).sum().item(), "absmax", cap["go"].abs().max().item())
for n in in_model:
    print(f"in-model vs oracle  {n:30s} {rel(in_model[n], P[n].grad):.2e}")
# standalone rerun on GPU with the same captured tensors
for p in mlp.parameters(): p.grad = None
hg = cap["h"].clone().requires_grad_(True)
y2 = mlp(hg); y2.backward(cap["go"])
for n, p in mlp.named_parameters():
    if p.grad is None: continue
    print(f"standalone vs oracle {n:30s} {rel(p.grad.cpu(), P[n].grad):.2e}")
