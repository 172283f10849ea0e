"""Where do the torch fill kernels of one QAT step come from? (torch profiler, python stacks)"""
import sys, collections
sys.path.insert(0, ".")
import torch, torch.nn.functional as F
import ofq_b200.quantization as Q
from ofq_b200.host.deit import deit_small_distilled_patch16_224
from ofq_b200.cga import CGAAdamW, param_groups_weight_decay
torch.manual_seed(0)
model = deit_small_distilled_patch16_224(num_classes=1000)
model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(Q.deit_qmodule_names(12), 2, 2), pretrained_initialized=True, qk_reparam=True).cuda()
img = torch.randn(16, 3, 224, 224, device="cuda"); lbl = torch.randint(0, 1000, (16,), device="cuda")
model.eval()
with torch.no_grad(): model(img)
model.train()
opt = CGAAdamW(param_groups_weight_decay(model, 0.05, model.no_weight_decay()), lr=1e-4)
def step():
    opt.zero_grad(set_to_none=True)
    (c, d), _ = model(img); loss = F.cross_entropy(c, lbl) + F.cross_entropy(d, lbl); loss.backward(); opt.step()
for _ in range(2): step()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
torch.cuda.synchronize()
cnt = collections.Counter(); ops_cnt = collections.Counter()
for e in prof.events():
    if e.name in ("aten::zeros", "aten::zero_", "aten::fill_", "aten::zeros_like", "aten::full", "aten::new_zeros"):
        st = [s for s in (e.stack or []) if "ofq_b200" in s or "torch/nn" in s or "autograd" in s][:2]
        cnt[(e.name, tuple(st))] += 1
    ops_cnt[e.name] += 1
for k, v in cnt.most_common(25): print(v, k)
print([ (k,v) for k,v in ops_cnt.most_common(40) if k.startswith("aten::")][:30])
