"""The drop-in boundary, functionally, on the GPU:

* a FOREIGN host model (ofq_b200/host/plain.py: timm-style classes that are not the repo's host classes, torch LayerNorm,
  un-fused residuals - what the reference's own src.deit_vision_transformer host looks like to the registry) gets its
  modules swapped by `replace_module_by_qmodule_deit`, loads a reference-generated state dict and reproduces the reference's
  golden logits / loss / gradients (the structural check against the reference's real classes is tests/test_reference_boundary.py);
* the CGA loop of cga.py:953-1013 restated verbatim (torch freeze masks, gradient masking, stash, torch.optim.AdamW, restore)
  runs UNCHANGED over the repo's modules, and the opt-in fused CGAAdamW produces the same weights: frozen elements bit-identical,
  trained ones to fp32 round-off."""
import copy

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from test_gpu_layers import check_grads, load_params, OUT_TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import ofq_b200.quantization as Q
    from ofq_b200 import _lib
    assert _lib.load().ofq_device_ok() == 1
    return Q


@pytest.mark.parametrize("qkr", [False, True])
def test_foreign_host_model_matches_reference_golden(Q, qkr):
    from ofq_b200.host.plain import PlainAttention, PlainDistilledViT
    g = load_golden(f"deit_tiny2_{'qkr' if qkr else 'plain'}_w2a2")
    model = PlainDistilledViT(embed_dim=64, depth=2, num_heads=2, num_classes=10)
    assert isinstance(model.blocks[0].attn, PlainAttention)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(Q.deit_qmodule_names(2), 2, 2), pretrained_initialized=True,
                                             qk_reparam=qkr)
    assert type(model.blocks[0].attn).__name__ == ("QAttention_qkreparam" if qkr else "QAttention")
    assert type(model.blocks[0]).__name__ == "PlainBlock" and type(model.blocks[0].norm1) is torch.nn.LayerNorm
    model = load_params(model, g).cuda().train()
    assert set(model.state_dict().keys()) == {k[len("param."):] for k in g if k.startswith("param.")}
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g["img_seed"]))).cuda()
    labels = g["labels"].cuda()
    (cls, dist), _ = model(img)
    assert rel_err(cls.detach().cpu(), g["cls"]) < OUT_TOL and rel_err(dist.detach().cpu(), g["dist"]) < OUT_TOL
    loss = F.cross_entropy(cls, labels) + F.cross_entropy(dist, labels)
    assert abs(loss.item() - g["loss"].item()) <= OUT_TOL * abs(g["loss"].item())
    loss.backward()
    assert check_grads(model.named_parameters(), g, sampled=True) > 50
    model.eval()
    with torch.no_grad():
        ev, _ = model(img)
    assert rel_err(ev.cpu(), g["eval_logits"]) < OUT_TOL


def _cga_loop_unchanged(model, optimizer, wq_bitw, boundary_range, freeze_fn):
    """cga.py:953-1013, DeiT + QKR branch, statement for statement (module-name suffix selection, torch masks, torch optimizer)."""
    save_frozen_weight, freeze_idx_dic = {}, {}
    for k, v in model.named_modules():
        if 'blocks' in k and (k[-3:] == 'fc1' or k[-3:] == 'fc2' or k[-2:] == '.v' or k[-4:] == 'proj'):
            freeze_idx = freeze_fn(v.weight, wq_bitw, boundary_range)
            freeze_idx_dic[k] = freeze_idx.detach().clone()
            v.weight.grad = v.weight.grad * freeze_idx * 0.0 + v.weight.grad * (1 - freeze_idx)
            save_frozen_weight[k] = (v.weight * freeze_idx).detach().clone()
    optimizer.step()
    for k, v in model.named_modules():
        if 'blocks' in k and (k[-3:] == 'fc1' or k[-3:] == 'fc2' or k[-2:] == '.v' or k[-4:] == 'proj'):
            with torch.no_grad():
                keep_w = (v.weight.detach().clone() * (1 - freeze_idx_dic[k]))
                new_weight = keep_w + save_frozen_weight[k]
                v.weight.data.copy_(new_weight)
    return freeze_idx_dic


def test_cga_loop_of_the_reference_runs_unchanged_and_fused_optimizer_agrees(Q):
    from oracle import ofq_oracle as O
    from ofq_b200.cga import CGAAdamW, cga_masked_parameter_names, param_groups_weight_decay
    from ofq_b200.host.deit import DistilledVisionTransformer
    torch.manual_seed(7)
    bits, br, lr, wd = 2, 0.05, 1e-3, 0.05
    model = DistilledVisionTransformer(embed_dim=64, depth=2, num_heads=2, num_classes=10)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(Q.deit_qmodule_names(2), bits, bits), pretrained_initialized=True,
                                             qk_reparam=True, qk_reparam_type=1, boundaryRange=br).cuda()
    img = torch.randn(4, 3, 224, 224, device="cuda")
    lbl = torch.tensor([1, 5, 7, 2], device="cuda")
    model.eval()
    with torch.no_grad():
        model(img)                                   # setup_alpha
    model.train()
    fused = copy.deepcopy(model)                     # second replica for the fused optimizer (own prologue: __deepcopy__)
    assert fused._ofq_prologue is not model._ofq_prologue
    groups = param_groups_weight_decay(model, wd, model.no_weight_decay())
    opt_ref = torch.optim.AdamW(groups, lr=lr)
    pd = dict(fused.named_parameters())
    masked = [pd[n] for n in cga_masked_parameter_names(fused, qk_reparam=True)]
    assert len(masked) == 2 * 4
    opt_fused = CGAAdamW(param_groups_weight_decay(fused, wd, fused.no_weight_decay()), lr=lr, masked=masked, wq_bitw=bits,
                         boundary_range=br)
    for step in range(3):
        before = {n: p.detach().clone() for n, p in model.named_parameters()}
        before_fused = {n: p.detach().clone() for n, p in fused.named_parameters()}
        # one forward / backward through the repo's modules; BOTH optimizers then see exactly these gradients (two replicas
        # running their own backward drift apart chaotically: split-K atomics, then flipped 2-bit codes)
        opt_ref.zero_grad(set_to_none=True)
        (cls, dst), _ = model(img)
        (F.cross_entropy(cls, lbl) + F.cross_entropy(dst, lbl)).backward()
        for (n, p), (_, q) in zip(model.named_parameters(), fused.named_parameters()):
            q.grad = None if p.grad is None else p.grad.detach().clone()
        masks = _cga_loop_unchanged(model, opt_ref, bits, br, O.cga_freeze_mask)
        opt_fused.step()
        nfrozen = 0
        for k, f in masks.items():
            w_new, w_old = dict(model.named_modules())[k].weight.detach(), before[k + ".weight"]
            frozen = f.bool()
            assert torch.equal(w_new[frozen], w_old[frozen]), k             # frozen weights restored bit-identically
            assert 0 < int((~frozen).sum()) < f.numel(), k                  # the band is neither empty nor everything
            nfrozen += int(frozen.sum())
        assert nfrozen > 0
        for (n, p), (_, q) in zip(model.named_parameters(), fused.named_parameters()):
            assert rel_err(q.detach(), p.detach()) < 1e-5, f"step {step} {n}: {rel_err(q.detach(), p.detach()):.2e}"
            if any(n == k + ".weight" for k in masks):
                # fused path: what ITS mask (the reference's formula on the replica's own weights) froze is untouched bit for bit
                frozen = O.cga_freeze_mask(before_fused[n], bits, br).bool()
                assert torch.equal(q.detach()[frozen], before_fused[n][frozen]), n
                assert bool((q.detach()[~frozen] != before_fused[n][~frozen]).any()), n


def test_cga_adamw_state_dict_round_trip_matches_torch_adamw(Q):
    """save -> load -> step must continue the bias correction (the kernels read a device step counter)."""
    from ofq_b200.cga import CGAAdamW
    torch.manual_seed(3)
    ps = [torch.nn.Parameter(torch.randn(33, 17, device="cuda")), torch.nn.Parameter(torch.randn(65, device="cuda"))]
    pr = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a, r = CGAAdamW(ps, lr=1e-2, weight_decay=0.05), torch.optim.AdamW(pr, lr=1e-2, weight_decay=0.05)
    gen = torch.Generator(device="cuda").manual_seed(5)

    def step(opts):
        gs = [torch.randn(p.shape, device="cuda", generator=gen) for p in ps]
        for opt, params in opts:
            for p, g_ in zip(params, gs):
                p.grad = g_.clone()
            opt.step()
    for _ in range(5):
        step(((a, ps), (r, pr)))
    sd = copy.deepcopy(a.state_dict())
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    b = CGAAdamW(ps2, lr=1e-2, weight_decay=0.05)
    b.load_state_dict(sd)
    for _ in range(3):
        gs = [torch.randn(p.shape, device="cuda", generator=gen) for p in ps]
        for opt, params in ((b, ps2), (r, pr)):
            for p, g_ in zip(params, gs):
                p.grad = g_.clone()
            opt.step()
    for p, q in zip(ps2, pr):
        assert rel_err(p.detach(), q.detach()) < 1e-6


def test_dispatcher_ops_run_the_c_abi_kernels():
    """torch.ops.ofq_b200.*  are thin wrappers over the same launchers as ofq_b200.ops: identical results."""
    import ofq_b200.torch_ops  # noqa: F401
    from ofq_b200 import ops
    torch.manual_seed(60)
    w = torch.nn.init.trunc_normal_(torch.empty(96, 64), std=0.02).cuda()
    codes, colscale = torch.ops.ofq_b200.statsq_codes(w, 2)
    rc, rs = ops.statsq_codes(w, 2)[:2]
    assert torch.equal(codes, rc) and torch.equal(colscale, rs)
    assert torch.equal(torch.ops.ofq_b200.cga_mask(w, 2, 0.05), ops.cga_mask(w, 2, 0.05))
    x = torch.randn(40, 64, device="cuda")
    b4 = torch.randn(64, device="cuda") * 0.05
    se = torch.ops.ofq_b200.lsq_effective_scale(torch.rand(10, device="cuda") * 0.5 + 0.2, 0.01)
    qx = torch.ops.ofq_b200.lsq_quant(x, b4, se, False, 10, 1, -2, 1)
    assert torch.equal(qx, ops.lsq_quant(x, b4, se, ops.PER_ROW, 10, 1, -2, 1))
    out = torch.ops.ofq_b200.qgemm_fwd(qx, se, codes, colscale, None)
    ref = (qx.float() * se.repeat(4).view(-1, 1)) @ (codes.float() * colscale.view(-1, 1)).t()
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    wq, wk = torch.randn(64, 64, device="cuda"), torch.randn(64, 64, device="cuda")
    assert torch.equal(torch.ops.ofq_b200.wqk_compose(wq, wk, 2), ops.wqk_compose(wq, wk, 2))
    p = torch.randn(96, 64, device="cuda")
    ref_p, g = p.clone(), torch.randn_like(p)
    m, v, rm, rv = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    torch.ops.ofq_b200.cga_adamw_step(p, g, m, v, 1, 1e-3, 0.9, 0.999, 1e-8, 0.05, 2, 0.05)
    ops.cga_adamw_(ref_p, g, rm, rv, 1, 1e-3, 0.9, 0.999, 1e-8, 0.05, bits=2, boundary_range=0.05)
    assert torch.equal(p, ref_p) and torch.equal(m, rm) and torch.equal(v, rv)
