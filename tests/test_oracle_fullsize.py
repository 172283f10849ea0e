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


@pytest.mark.parametrize("cfg", [c for c in FC.CONFIGS if FC.CONFIGS[c][4] == 0])
def test_exact_gemm_oracle_differs_from_the_reference_only_by_proven_ties(cfg):
    """The oracle's EXACT_GEMM mode (matrix products accumulated in float64, rounded once) is what the CUDA path is compared
    with on the GPU. Here it is pinned to the reference arithmetic (fp32 sgemm mode, bit-identical to the reference): every
    block of the model is evaluated in both modes ON THE SAME block input (taken from the fp32 run); at every quantizer
    that is not downstream of an already flipped code of that block, the codes must agree except for PROVEN TIES (pre-round
    values within 2e-5 * max(1, |v|) of each other, i.e. the reference's own sgemm round-off decided the rounding)."""
    torch.set_num_threads(8)
    g = load_golden(f"full_{cfg}")
    model_name = FC.CONFIGS[cfg][0]
    P = FC.oracle_params(cfg, g, requires_grad=False)
    img = FC.det_images()
    O.TAPS = {}
    try:
        with torch.no_grad():
            FC.oracle_forward(cfg, P, img, int(g["signed"]))
        free = O.TAPS
    finally:
        O.TAPS = None
    prefixes = FC.block_prefixes(model_name)
    total_sites = flips = contaminated_sites = 0
    worst_clean_out = 0.0
    for bi, pre in enumerate(prefixes):
        x_in = free[pre + "@in"]["x"]
        res = {}
        for exact in (False, True):
            O.TAPS, O.EXACT_GEMM = {}, exact
            try:
                with torch.no_grad():
                    out = FC.oracle_block(cfg, P, pre, x_in, bi)
                res[exact] = (out, O.TAPS)
            finally:
                O.TAPS, O.EXACT_GEMM = None, False
        dirty = 0
        # the only weight whose codes depend on a matrix product: StatsQ of W_q^T W_k (QKR). A flipped WEIGHT code changes a
        # whole column of qkx: it must itself be a proven tie, and everything after it in the block is downstream of it
        wq_flips = 0
        if pre + "attn.q.weight" in P:
            codes = {}
            for exact in (False, True):
                O.EXACT_GEMM = exact
                try:
                    w = FC.weight_of_site(P, pre, "attn.qk_quant")
                finally:
                    O.EXACT_GEMM = False
                codes[exact] = (O.statsq_codes(w, FC.CONFIGS[cfg][1])[0], O.statsq_pre_round(w, FC.CONFIGS[cfg][1])[0])
            bad = codes[True][0] != codes[False][0]
            wq_flips = int(bad.sum())
            if wq_flips:
                assert bool(((codes[True][1][bad] - codes[False][1][bad]).abs() <= FC.TIE_TOL).all()), (cfg, bi, "qk_quant")
            flips += wq_flips
        for (name, a), (_, b) in zip(FC.lsq_sites(res[False][1], pre), FC.lsq_sites(res[True][1], pre)):
            if name.endswith("quan_a_qkx_fn"):
                dirty += wq_flips
            r = FC.compare_codes(b["codes"], b["x"] / b["se"], a["codes"], a["x"] / a["se"])
            total_sites += 1
            if dirty == 0:
                assert r["not_ties"] == 0, (cfg, bi, name, r)
            else:
                contaminated_sites += 1
            dirty += r["mismatches"]
            flips += r["mismatches"]
        if dirty == 0:
            worst_clean_out = max(worst_clean_out, rel_err(res[True][0], res[False][0]))
    print(f"{cfg}: {flips} flipped codes in {total_sites} quantizer calls ({contaminated_sites} downstream of a flip); "
          f"blocks without a flip agree to {worst_clean_out:.1e}")
    assert worst_clean_out < 1e-5
