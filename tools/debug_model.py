import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, torch.nn as nn, torch.nn.functional as F
import ofq_b200.quantization as Q
from ofq_b200.host.deit import DistilledVisionTransformer
from oracle import ofq_oracle as O
from conftest import rel_err, load_golden
g = load_golden("deit_tiny2_plain_w2a2")
model = DistilledVisionTransformer(embed_dim=64, depth=2, num_heads=2, num_classes=10)
names = Q.deit_qmodule_names(2)
model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(names, 2, 2), pretrained_initialized=True, qk_reparam=False)
sd = {k[6:]: v for k, v in g.items() if k.startswith("param.")}
print(model.load_state_dict(sd, strict=False))
model = model.cuda().train()
img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["img_seed"])))
P = {k: v.clone() for k, v in sd.items()}
state = {"signed": 1}
with torch.no_grad():
    xo = O.patch_embed_q(img, P, "patch_embed.proj.", state).flatten(2).transpose(1, 2)
    xg = model.patch_embed(img.cuda())
    print("patch", rel_err(xg.cpu(), xo))
    xo = torch.cat((P["cls_token"].expand(2, -1, -1), P["dist_token"].expand(2, -1, -1), xo), 1) + P["pos_embed"]
    xg = torch.cat((model.cls_token.expand(2, -1, -1), model.dist_token.expand(2, -1, -1), xg), 1) + model.pos_embed
    for i, blk in enumerate(model.blocks):
        pre = f"blocks.{i}."
        ho = F.layer_norm(xo, (64,), P[pre + "norm1.weight"], P[pre + "norm1.bias"], 1e-6)
        hg = blk.norm1(xg)
        print(i, "norm1", rel_err(hg.cpu(), ho))
        qkvo = O.qlinear(ho, P, pre + "attn.qkv.", 2, 2)
        qkvg = blk.attn.qkv(hg)
        print(i, "qkv", rel_err(qkvg.cpu(), qkvo))
        ao = O.qattention(ho, P, pre + "attn.", 2, 2, 2)
        ag, _ = blk.attn(hg)
        print(i, "attn", rel_err(ag.cpu(), ao))
        # feed the oracle's input to the GPU module to isolate
        ag2, _ = blk.attn(ho.cuda())
        print(i, "attn (oracle input)", rel_err(ag2.cpu(), ao))
        xo = xo + ao; xg = xg + ag
        ho = F.layer_norm(xo, (64,), P[pre + "norm2.weight"], P[pre + "norm2.bias"], 1e-6)
        hg = blk.norm2(xg)
        mo = O.qmlp(ho, P, pre + "mlp.", 2, 2)
        mg = blk.mlp(hg)
        print(i, "mlp", rel_err(mg.cpu(), mo))
        xo = xo + mo; xg = xg + mg
