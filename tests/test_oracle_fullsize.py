"""Pin oracle/ofq_oracle.py at the REAL width and depth of every BASELINE.json configuration against golden outputs of the
unmodified reference (tests/golden/make_golden_fullsize.py): depth-12 DeiT-T / DeiT-S (plain and QKR, type 0 and 1) and
Swin-T with its real stage dimensions, batch 8. Forward: logits, loss, every block output (sampled) and every integer code
tensor of the first and last block must be identical (same fp32 op sequence); gradients (sampled) <= 2e-5.

The GPU parity tests (tests/test_gpu_fullsize_parity.py) then compare the CUDA path with this oracle run live on the same
regenerated parameters, so the chain is: reference == oracle (here, full size) == CUDA path (GPU box, full size)."""
import pytest
import torch
import torch.nn.functional as F

import fullsize_common as FC
from conftest import load_golden, rel_err
from oracle import ofq_oracle as O


@pytest.mark.parametrize("cfg", list(FC.CONFIGS))
def test_oracle_matches_reference_fullsize(cfg):
    torch.set_num_threads(8)
    g = load_golden(f"full_{cfg}")
    model_name = FC.CONFIGS[cfg][0]
    P = FC.oracle_params(cfg, g)
    img, labels = FC.det_images(), FC.det_labels(FC.BATCH, 1000)
    O.TAPS = {}
    try:
        outs = FC.oracle_forward(cfg, P, img, int(g["signed"]))
        taps = O.TAPS
    finally:
        O.TAPS = None
    loss = sum(F.cross_entropy(o, labels) for o in outs)
    # ---- forward: identical op sequence in fp32 -> identical values (a last-bit slack keeps the test portable across BLAS builds)
    ref_logits = (g["logits"],) if model_name == "swin_tiny" else (g["cls"], g["dist"])
    for o, r in zip(outs, ref_logits):
        assert rel_err(o.detach(), r) < 1e-6
    assert abs(loss.item() - g["loss"].item()) <= 1e-6 * abs(g["loss"].item())
    prefixes = FC.block_prefixes(model_name)
    for bi, pre in enumerate(prefixes):
        mine = FC.sample(taps[pre + "@out"]["out"], 4096)
        assert rel_err(mine, g[f"block{bi}.out_sample"]) < 1e-6, f"block {bi}"
    # ---- integer codes of the first and the last block
    nimg = int(g["code_images"])
    checked = 0
    for tag, pre in (("first", prefixes[0]), ("last", prefixes[-1])):
        for k, ref in g.items():
            if not k.startswith(f"{tag}.codes."):
                continue
            t = taps[pre + k[len(f"{tag}.codes."):]]["codes"]
            per = t.shape[0] // FC.BATCH
            mine = t[: nimg * per].reshape(ref.shape)
            nbad = int((mine != ref).sum())
            assert nbad <= 1e-5 * ref.numel(), f"{k}: {nbad} of {ref.numel()} codes differ"
            checked += 1
        for k, ref in g.items():
            if not k.startswith(f"{tag}.wcodes."):
                continue
            rel = k[len(f"{tag}.wcodes."):]
            # weight codes from the oracle's StatsQ on the same weights
            if rel.endswith("qk_quant"):
                a = pre + rel[: -len("qk_quant")]
                heads = P[a + "quan_a_qkx_fn.s"].numel() // P[a + "quant_x_4_qkv.input_quant_fn.s"].numel()
                w = O.wqk_compose(P[a + "q.weight"].detach(), P[a + "k.weight"].detach(), heads)
            elif rel.endswith("v_quant"):
                w = P[pre + rel[: -len("v_quant")] + "v.weight"].detach()
            else:
                w = P[pre + rel[: -len("statsq_fn")] + "weight"].detach()
            codes, _ = O.statsq_codes(w, FC.CONFIGS[cfg][1])
            mine = FC.row_sample(codes.to(torch.int8))
            assert int((mine != ref).sum()) <= 1e-5 * ref.numel(), k
            checked += 1
    assert checked > 0 or nimg == 0
    # ---- gradients (strided samples of every parameter gradient)
    loss.backward()
    n = 0
    gmax = max(v.abs().max().item() for k, v in g.items() if k.startswith("grad."))
    for k, ref in g.items():
        if not k.startswith("grad."):
            continue
        p = P[k[len("grad."):]]
        assert p.grad is not None, k
        mine = FC.sample(p.grad, 2048)
        assert rel_err(mine, ref) < 2e-5 or (mine - ref).abs().max().item() <= 1e-6 * gmax, f"{k}: {rel_err(mine, ref):.2e}"
        n += 1
    assert n > 100
