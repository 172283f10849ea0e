"""GPU probe: run every golden layer / model case through the ofq_b200 modules and print relative errors."""
import sys
from functools import partial
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, torch.nn as nn
import torch.nn.functional as F
from conftest import load_golden, rel_err
import ofq_b200.quantization as Q
from ofq_b200.host.deit import Attention, Mlp, DistilledVisionTransformer

torch.manual_seed(0)
dev = "cuda"

def load_params(mod, g):
    sd = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    missing, unexpected = mod.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not [m for m in missing if not m.endswith(".s")], missing
    return mod

def run(name, mod, g):
    mod = load_params(mod, g).to(dev).train()
    x = g["x"].to(dev).requires_grad_(True)
    y = mod(x)
    y = y[0] if isinstance(y, tuple) else y
    print(f"{name}: out rel {rel_err(y.cpu(), g['out']):.2e}", end="")
    y.backward(g["go"].to(dev))
    print(f"  dx {rel_err(x.grad.cpu(), g['dx']):.2e}")
    worst = []
    for k, v in g.items():
        if k.startswith("grad."):
            p = dict(mod.named_parameters())[k[5:]]
            if p.grad is None:
                print("   MISSING grad", k); continue
            e = (p.grad.cpu() - v).norm().item() / max(v.norm().item(), 1e-30)
            worst.append((e, k[5:], v.norm().item()))
    worst.sort(reverse=True)
    for e, k, n in worst[:8]:
        print(f"   {k:45s} rel {e:.2e}  |ref| {n:.2e}")

for bits in (2, 4):
    C, H = 32, 2
    run(f"qlinear w{bits}", Q.QLinear(m=nn.Linear(C, 48), weight_bits=bits, input_bits=bits), load_golden(f"qlinear_w{bits}a{bits}"))
    run(f"qmlp w{bits}", Q.QMLP(m=Mlp(C, 4 * C), weight_bits=bits, input_bits=bits), load_golden(f"qmlp_w{bits}a{bits}"))
    run(f"qattn w{bits}", Q.QAttention(Attention(C, H, qkv_bias=True), weight_bits=bits, input_bits=bits), load_golden(f"qattention_w{bits}a{bits}"))
    run(f"qattn_qkr w{bits}", Q.QAttention_qkreparam(Attention(C, H, qkv_bias=True), weight_bits=bits, input_bits=bits), load_golden(f"qattention_qkr_w{bits}a{bits}"))

for qkr in (False, True):
    g = load_golden(f"deit_tiny2_{'qkr' if qkr else 'plain'}_w2a2")
    model = DistilledVisionTransformer(embed_dim=64, depth=2, num_heads=2, num_classes=10)
    names = Q.deit_qmodule_names(2)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(names, 2, 2), pretrained_initialized=True, qk_reparam=qkr)
    model = load_params(model, g).to(dev).train()
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["img_seed"]))).to(dev)
    (cls, dist), _ = model(img)
    loss = F.cross_entropy(cls, g["labels"].to(dev)) + F.cross_entropy(dist, g["labels"].to(dev))
    print(f"deit qkr={qkr}: cls rel {rel_err(cls.cpu(), g['cls']):.2e} dist {rel_err(dist.cpu(), g['dist']):.2e} loss {loss.item():.6f} vs {g['loss'].item():.6f}")
    loss.backward()
    worst = []
    for n, p in model.named_parameters():
        k = "gnorm." + n
        if k not in g: continue
        if p.grad is None:
            print("   MISSING", n); continue
        gr = p.grad.cpu()
        ref = g["grad." + n]
        mine = gr if gr.numel() <= 4096 else gr.flatten()[:: max(1, gr.numel() // 2048)][:2048]
        e = (mine - ref).norm().item() / max(ref.norm().item(), 1e-30)
        worst.append((e, n, g[k].item()))
    worst.sort(reverse=True)
    for e, k, n in worst[:12]:
        print(f"   {k:55s} rel {e:.2e}  |ref| {n:.2e}")
    import statistics
    print("   median grad rel err", statistics.median(w[0] for w in worst))
    model.eval()
    with torch.no_grad():
        ev, _ = model(img)
    print(f"   eval logits rel {rel_err(ev.cpu(), g['eval_logits']):.2e}")
