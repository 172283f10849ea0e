"""Deployment export: packed low-bit weight codes and an inference forward that runs from them alone."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("shape", [(384, 384), (1536, 384), (7, 13), (96, 100)])
def test_pack_unpack_round_trip_is_exact(bits, shape):
    from ofq_b200 import _lib, export
    assert _lib.load().ofq_device_ok() == 1
    torch.manual_seed(bits)
    n = 1 << (bits - 1)
    k = torch.randint(-n, n, shape, device="cuda")
    codes = (2 * k + 1).to(torch.int8)
    packed = export.pack_codes(codes, bits)
    assert packed.shape == (shape[0], (shape[1] + 7) // 8 * bits)                   # true width (rows padded to 8 codes)
    assert torch.equal(export.unpack_codes(packed, shape[1], bits), codes)
    # the layout is the documented one: code e of a group in bits [e*b, (e+1)*b) of the group's little-endian word
    u = (k[0, :min(8, shape[1])] + n).tolist()
    word = sum(v << (e * bits) for e, v in enumerate(u))
    assert packed[0, :bits].tolist() == [(word >> (8 * b)) & 255 for b in range(bits)]


@pytest.mark.parametrize("qkr", [True, False])
def test_packed_model_runs_from_the_codes_alone(qkr):
    """export_packed -> load_packed into a FRESH model whose quantized fp32 weights are then zeroed: the inference logits are
    bit-identical to the trained model's, the weight-side kernels (StatsQ, W_qk compose) no longer run, and the packed codes
    take bits/32 of the fp32 weights they replace."""
    import ofq_b200.quantization as Q
    from ofq_b200 import export, ops
    from ofq_b200.host.deit import DistilledVisionTransformer
    depth = 2

    def build(seed):
        torch.manual_seed(seed)
        m = DistilledVisionTransformer(embed_dim=128, depth=depth, num_heads=2, num_classes=10)
        return Q.replace_module_by_qmodule_deit(m, Q.make_qconfigs(Q.deit_qmodule_names(depth), 2, 2), pretrained_initialized=True,
                                                qk_reparam=qkr).cuda()

    img = torch.randn(3, 3, 224, 224, device="cuda")
    model = build(50)
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if n_.endswith(".bias") and p.dim() == 1:
                p.normal_(0, 0.02)
    model.eval()
    with torch.no_grad():
        ref = model(img)[0].clone()
    packed = export.export_packed(model, img)
    sizes = export.packed_nbytes(packed)
    assert sizes["codes"] * 8 == sizes["quantized_weights"] * 2                      # 2 bits per weight
    assert len(packed["sites"]) == depth * (5 if qkr else 4)                        # QKR: W_qk, v, proj, fc1, fc2 / plain: qkv, proj, fc1, fc2
    if qkr:
        assert any(k.endswith(".W_qk") for k in packed["sites"]) and all(not k.endswith(("q.weight", "k.weight")) for k in packed["state"])
    fresh = build(51)                                                               # other random weights
    fresh = export.load_packed(fresh, packed, img, drop_fp32=True)
    params = dict(fresh.named_parameters())
    assert all(float(params[n].detach().abs().max()) == 0.0 for n in packed["replaced"])      # the fp32 weights are gone
    ops.PROFILE = []
    try:
        with torch.no_grad():
            out = fresh(img)[0]
        fams = [r[0] for r in ops.PROFILE]
    finally:
        ops.PROFILE = None
    assert torch.equal(out, ref)
    assert "statsq" not in fams and "wqk_compose" not in fams
