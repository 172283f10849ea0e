import sys
sys.path.insert(0, ".")
import torch, torch.nn.functional as F
import ofq_b200.quantization as Q
from ofq_b200.host.deit import DistilledVisionTransformer
from oracle import ofq_oracle as O
torch.manual_seed(0)
depth, dim, heads = 2, 64, 2
model = DistilledVisionTransformer(embed_dim=dim, depth=depth, num_heads=heads, num_classes=10)
names = Q.deit_qmodule_names(depth)
model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(names, 2, 2), pretrained_initialized=True, qk_reparam=True, qk_reparam_type=1).cuda()
img = torch.randn(2, 3, 224, 224); labels = torch.tensor([1, 5])
model.eval()
with torch.no_grad(): model(img.cuda())
model.train()
(cls, dist), _ = model(img.cuda())
loss = F.cross_entropy(cls, labels.cuda()) + F.cross_entropy(dist, labels.cuda()); loss.backward()
P = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
state = {"signed": int(P["patch_embed.proj.input_quant_fn.signed"].item())}
ocls, odist = O.deit_forward(img, P, depth, heads, 2, 2, qkr=True, state=state)
oloss = F.cross_entropy(ocls, labels) + F.cross_entropy(odist, labels); oloss.backward()
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
print("cls rel", rel(cls.detach().cpu(), ocls.detach()), "loss", loss.item(), oloss.item())
errs = sorted(((rel(p.grad.cpu(), P[n].grad), n, P[n].grad.norm().item()) for n, p in model.named_parameters() if P[n].grad is not None), reverse=True)
for e in errs[:15]: print(f"{e[1]:50s} {e[0]:.2e} |ref| {e[2]:.2e}")
n = 'blocks.0.mlp.fc1.input_quant_fn.s'
a = dict(model.named_parameters())[n].grad.cpu(); b = P[n].grad
d = (a - b).abs(); print("bad entries:", (d > 1e-3 * b.abs().max()).sum().item(), "of", d.numel(), d.topk(5))
print(a[:8], b[:8])
