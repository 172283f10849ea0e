"""GPU parity at the REAL width and depth of every BASELINE.json configuration (depth-12 DeiT-T / DeiT-S, Swin-T stage
dimensions), against (a) golden outputs of the unmodified reference (tests/golden/full_*.npz, batch 8) and (b) the CPU oracle
run live on the same regenerated parameters (pinned to those goldens bit for bit by tests/test_oracle_fullsize.py).

Low-bit QAT is chaotic in the last bit: ONE 2-bit code that sits on a rounding tie and flips changes its block's output by
~1 % and the logits of a depth-12 model by tens of percent (measured on the reference arithmetic itself: a 1-ulp perturbation
of one LayerNorm bias moves Swin-T's logits by 6 %). The reference's fp32 sgemm carries ~1e-6 of accumulated rounding error;
the CUDA path's integer GEMMs are exact. test_oracle_fullsize.py therefore pins a second oracle mode, EXACT_GEMM (products
accumulated in float64, rounded once), to the reference arithmetic: the two modes differ ONLY by proven ties. The tests here:

* teacher forced, EVERY block (each block is fed the reference's own block input, first two images): all integer code
  tensors (activation codes of every quantizer, StatsQ weight codes) must equal the exact-GEMM oracle's bit for bit; a
  mismatch is accepted only as a PROVEN TIE (pre-round values of both sides within 2e-5 * max(1, |v|) of each other).
  Block output <= 1e-5 and EVERY gradient of the block (d input, weights, biases, shifts, LSQ step sizes) <= 1e-3 relative,
  asserted for every block in which no tie flipped (a flip is reported; its block is contaminated by construction).
* free running (whole model, batch 8, reference goldens): logits / loss / per-block outputs / gradients / codes of the first
  and last block are REPORTED per block; asserted: the divergence from the reference is no larger than the divergence of the
  reference arithmetic from itself under exact GEMMs (same kind of difference, measured live), the loss agrees to 2 %.

A JSON report per configuration is written to gpurun_out/fullsize_parity_<cfg>.json (summarised in profiles/)."""
import json
import os
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

import fullsize_common as FC
from conftest import ROOT, load_golden, rel_err

pytestmark = pytest.mark.gpu
OUT_TOL, GRAD_TOL = 1e-3, 1e-3
TIE_TOL = FC.TIE_TOL
ANALYTIC_ZERO = ("move_k_aft.bias", "move_qkx_aft.bias")
REPORT_DIR = ROOT / "gpurun_out"


@pytest.fixture(scope="module")
def Fn():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from ofq_b200 import _lib
    from ofq_b200.quantization import functional as Fn
    assert _lib.load().ofq_device_ok() == 1
    return Fn


def _report(cfg, section, payload):
    REPORT_DIR.mkdir(exist_ok=True)
    path = REPORT_DIR / f"fullsize_parity_{cfg}.json"
    data = json.loads(path.read_text()) if path.exists() else {}
    data[section] = payload
    path.write_text(json.dumps(data, indent=1, sort_keys=True))


# ------------------------------------------------------------------------------------------------ our codes / pre-round values
def _rows(se, M, period):
    return se.repeat(M // period).unsqueeze(1)


def _sites_of_block(taps, qkr, swin=False):
    """[(reference module name relative to the block, codes int tensor, pre-round fp32 tensor)] in forward order, laid out as
    the reference lays its tensors out (so that they compare element for element with the oracle's taps)."""
    out = []
    i = 0
    if qkr:
        kind, t = taps[i]; i += 1
        assert kind == "qkr"
        B, N, H, C = t["B"], t["N"], t["H"], t["C"]
        M = B * N
        out.append(("attn.quant_x_4_qkv.input_quant_fn", t["qx"].view(B, N, C), ((t["x"] + t["x_b4"]) / _rows(t["se_x"], M, N)).view(B, N, C)))
        out.append(("attn.quan_a_v_fn", t["qv"].view(B, N, C), ((t["v_out"] + t["v_b4"]) / t["se_v"].view(1, C)).view(B, N, C)))
        sek = t["se_k"].view(1, N, H, 1)
        out.append(("attn.quan_a_qkx_fn", t["qk"].view(B, N * H, C),
                    ((t["qkx"].view(B, N, H, C) + t["k_b4"].view(1, 1, H, C)) / sek).reshape(B, N * H, C)))
        Pm = t["P"][..., :N].reshape(B, H, N, N)
        out.append(("attn.quan_a_softmax_fn", t["qp"][..., :N].reshape(B, H, N, N), Pm / t["se_p"].view(1, 1, N, 1)))
    else:
        kind, t = taps[i]; i += 1
        assert kind == "qlinear"
        out.append(("attn.qkv.input_quant_fn",) + _qlinear_site(t))
        kind, t = taps[i]; i += 1
        assert kind == "attn"
        B, N, H, C = t["B"], t["N"], t["H"], t["C"]
        M, hd = B * N, t["C"] // t["H"]
        x = t["qkv"] + t["b4"].view(1, 3 * C)

        def heads(a):      # [M, C] (h*hd+j) -> (B, H, N, hd)
            return a.reshape(B, N, H, hd).permute(0, 2, 1, 3)
        out.append(("attn.quan_a_q_fn", heads(t["qq"]), heads(x[:, :C] / _rows(t["se_q"], M, N))))
        out.append(("attn.quan_a_k_fn", heads(t["qk"]), heads(x[:, C:2 * C] / _rows(t["se_k"], M, N))))
        out.append(("attn.quan_a_v_fn", t["qv"].view(B, N, C), (x[:, 2 * C:] / t["se_v"].view(1, C)).view(B, N, C)))
        Pm = t["P"][..., :N].reshape(B, H, N, N)
        out.append(("attn.quan_a_softmax_fn", t["qp"][..., :N].reshape(B, H, N, N), Pm / t["se_p"].view(1, 1, N, 1)))
    for name in ("attn.proj.input_quant_fn", "mlp.fc1.input_quant_fn", "mlp.fc2.input_quant_fn"):
        kind, t = taps[i]; i += 1
        assert kind == "qlinear", (name, kind)
        out.append((name,) + _qlinear_site(t))
    return out, i


def _qlinear_site(t):
    x = t["x"]
    if t["act"] == 1:
        x = F.gelu(x)
    M = x.shape[0]
    return t["qx"], (x + t["b4"]) / _rows(t["se"], M, t["period"])


def _weight_sites(taps, qkr):
    """[(reference quantizer name, our odd weight codes [rows, cols])] of one block."""
    out = []
    i = 0
    if qkr:
        t = taps[i][1]; i += 1
        out += [("attn.v_quant", t["wvc"]), ("attn.qk_quant", t["wqkc"])]
    else:
        out.append(("attn.qkv.statsq_fn", taps[i][1]["wc"])); i += 2
    for name in ("attn.proj.statsq_fn", "mlp.fc1.statsq_fn", "mlp.fc2.statsq_fn"):
        out.append((name, taps[i][1]["wc"])); i += 1
    return out


_compare_codes = FC.compare_codes


# ------------------------------------------------------------------------------------------------ free running
def _block_modules(model, model_name):
    if model_name == "swin_tiny":
        return [model.features[2 * i + 1][j] for i, d in enumerate((2, 2, 6, 2)) for j in range(d)]
    return list(model.blocks)


@pytest.mark.parametrize("cfg", list(FC.CONFIGS))
def test_free_running_against_reference(Fn, cfg):
    g = load_golden(f"full_{cfg}")
    model_name, wb, ab, qkr, _, nimg = FC.CONFIGS[cfg]
    swin = model_name == "swin_tiny"
    model = FC.load_repo_model(cfg, g).cuda().train()
    img, labels = FC.det_images().cuda(), FC.det_labels(FC.BATCH, 1000).cuda()
    blocks = _block_modules(model, model_name)
    outs = {}
    hooks = []
    for bi, blk in enumerate(blocks):
        def hook(m, inp, out, bi=bi):
            if isinstance(out, tuple):          # (x, pending MLP branch): the block output is their sum (host/deit.py)
                out = out[0] if out[1] is None else out[0] + out[1]
            outs[bi] = FC.sample(out, 4096).float().cpu()
        hooks.append(blk.register_forward_hook(hook))
    Fn.TAP = []
    try:
        res = model(img)[0]
    finally:
        taps, Fn.TAP = Fn.TAP, None
        for h in hooks:
            h.remove()
    logits = (res,) if swin else res
    refs = (g["logits"],) if swin else (g["cls"], g["dist"])
    rep = {"logits_rel_err": [rel_err(a.detach().cpu(), r) for a, r in zip(logits, refs)]}
    loss = sum(F.cross_entropy(o, labels) for o in logits)
    rep["loss"], rep["loss_ref"] = loss.item(), g["loss"].item()
    rep["block_out_rel_err"] = [rel_err(outs[bi], g[f"block{bi}.out_sample"]) for bi in range(len(blocks))]
    # ---- codes of the first / last block against the reference's (free running: informative for the last block)
    if nimg > 0:
        per_block = len(taps) // len(blocks) if not swin else None
        code_rep = {}
        for tag, bi in (("first", 0), ("last", len(blocks) - 1)):
            if swin:        # reduction QLinears sit between stages: locate the block's taps by counting
                start = _swin_tap_start(bi, qkr)
            else:
                start = bi * per_block
            sites, used = _sites_of_block(taps[start:], qkr, swin)
            for name, codes, _ in sites:
                ref = g[f"{tag}.codes.{name}"]
                per = codes.shape[0] // FC.BATCH
                code_rep[f"{tag}.{name}"] = _compare_codes(codes[: nimg * per], None, ref)
            for name, wc in _weight_sites(taps[start:], qkr):
                ref = g[f"{tag}.wcodes.{name}"]
                code_rep[f"{tag}.w.{name}"] = _compare_codes(FC.row_sample(wc), None, ref)
        rep["codes_vs_reference"] = code_rep
    # ---- gradients
    loss.backward()
    gmax = max(v.abs().max().item() for k, v in g.items() if k.startswith("grad."))
    params = dict(model.named_parameters())
    worst, failures, per_block_worst = 0.0, [], {}
    for k, ref in g.items():
        if not k.startswith("grad."):
            continue
        name = k[len("grad."):]
        p = params[name]
        assert p.grad is not None, name
        mine = FC.sample(p.grad, 2048).float().cpu()
        if name.endswith(ANALYTIC_ZERO):
            if mine.abs().max().item() > 1e-5 * gmax:
                failures.append((name, "analytic zero", mine.abs().max().item() / gmax))
            continue
        e = rel_err(mine, ref)
        small = (mine - ref).abs().max().item() <= 1e-5 * gmax
        if not small:
            worst = max(worst, e)
            key = ".".join(name.split(".")[:3 if swin else 2])
            per_block_worst[key] = max(per_block_worst.get(key, 0.0), e)
        if not (e < GRAD_TOL or small):
            failures.append((name, e))
    rep["grad_worst_rel_err"] = worst
    rep["grad_worst_per_block"] = per_block_worst
    rep["grad_failures_free_running"] = len(failures)
    model.eval()
    with torch.no_grad():
        ev = model(img)[0]
    rep["eval_logits_rel_err"] = rel_err(ev.cpu(), g["eval_logits"])
    from ofq_b200.quantization.functional import BWD_MODE
    # ---- the reference arithmetic against itself: exact-GEMM oracle vs the reference goldens (same images, same parameters)
    from oracle import ofq_oracle as O
    P = FC.oracle_params(cfg, g, requires_grad=False)
    O.EXACT_GEMM = True
    try:
        with torch.no_grad():
            ex = FC.oracle_forward(cfg, P, FC.det_images(), int(g["signed"]))
    finally:
        O.EXACT_GEMM = False
    rep["reference_self_divergence_logits"] = [rel_err(a_, r) for a_, r in zip(ex, refs)]
    rep["logits_rel_err_vs_exact_oracle"] = [rel_err(a_.detach().cpu(), b_) for a_, b_ in zip(logits, ex)]
    _report(cfg, f"free_running[{BWD_MODE}]", rep)
    bound = max(1e-3, 3.0 * max(rep["reference_self_divergence_logits"]))
    assert max(rep["logits_rel_err"]) <= bound or max(rep["logits_rel_err_vs_exact_oracle"]) < OUT_TOL, rep
    assert abs(rep["loss"] - rep["loss_ref"]) <= 2e-2 * abs(rep["loss_ref"])
    assert rep["block_out_rel_err"][0] < 5e-2, rep["block_out_rel_err"]


def _swin_tap_start(bi, qkr):
    """Index of the first tap of Swin block `bi` (12 blocks over stages of depth 2, 2, 6, 2; one `reduction` QLinear tap
    after each of the first three stages)."""
    per = 4 if qkr else 5
    stage_end = (2, 4, 10, 12)
    reductions = sum(1 for e in stage_end[:3] if bi >= e)
    return bi * per + reductions


# ------------------------------------------------------------------------------------------------ teacher forced
def _weight_flips(taps, qkr, P, pre, wb, rep, key):
    """Compare our StatsQ weight codes of one block with the oracle's; returns {reference quantizer name: flipped codes}.
    A flip must be a proven tie: its pre-round value lies within TIE_TOL of the rounding boundary."""
    from oracle import ofq_oracle as O
    flips = {}
    for name, wc in _weight_sites(taps, qkr):
        w = FC.weight_of_site(P, pre, name)
        ref_codes, _ = O.statsq_codes(w, wb)
        r = _compare_codes(wc, None, ref_codes)
        if r["mismatches"]:
            b4, _ = O.statsq_pre_round(w, wb)
            bad = wc.detach().cpu().to(torch.int32) != ref_codes
            frac = (b4[bad] - torch.floor(b4[bad]) - 0.5).abs()
            r["ties"] = int((frac <= TIE_TOL).sum())
            r["not_ties"] = r["mismatches"] - r["ties"]
            rep[f"{key}.w.{name}"] = r
        flips[name] = r["mismatches"]
        assert r.get("not_ties", 0) == 0, (key, name, r)
    return flips


@pytest.mark.parametrize("cfg", [c for c in FC.CONFIGS if FC.CONFIGS[c][4] == 0])
def test_teacher_forced_every_block_against_exact_oracle(Fn, cfg):
    from oracle import ofq_oracle as O
    g = load_golden(f"full_{cfg}")
    model_name, wb, ab, qkr, _, _ = FC.CONFIGS[cfg]
    swin = model_name == "swin_tiny"
    nimg = 2
    model = FC.load_repo_model(cfg, g).cuda().train()
    P = FC.oracle_params(cfg, g, requires_grad=True)
    blocks = _block_modules(model, model_name)
    prefixes = FC.block_prefixes(model_name)
    # the reference's own activations: block inputs of the fp32 (reference-identical) free run, batch 8
    O.TAPS = {}
    try:
        with torch.no_grad():
            FC.oracle_forward(cfg, P, FC.det_images(), int(g["signed"]))
        free = O.TAPS
    finally:
        O.TAPS = None
    rep = {"blocks_clean": 0, "blocks_with_flipped_tie": 0, "worst_clean_block_out": 0.0, "worst_clean_grad": 0.0,
           "worst_clean_grad_name": "", "flipped_ties": 0}
    for bi, (blk, pre) in enumerate(zip(blocks, prefixes)):
        x_in = free[pre + "@in"]["x"][:nimg].clone()
        gy = FC.det_normal(tuple(x_in.shape), 9000 + bi, 1e-3)
        # ---- exact-GEMM oracle: forward taps + autograd gradients of this block alone
        bp = {k: v for k, v in P.items() if k.startswith(pre)}
        for v in bp.values():
            v.grad = None
        xo = x_in.clone().requires_grad_(True)
        O.TAPS, O.EXACT_GEMM = {}, True
        try:
            out_ref = FC.oracle_block(cfg, P, pre, xo, bi)
            otaps = O.TAPS
        finally:
            O.TAPS, O.EXACT_GEMM = None, False
        out_ref.backward(gy)
        # ---- CUDA path
        blk.zero_grad(set_to_none=True)
        xg = x_in.cuda().requires_grad_(True)
        Fn.TAP = []
        try:
            if swin:
                out = blk(xg)
            else:
                blk._defer = False
                out, _ = blk(xg)
        finally:
            taps, Fn.TAP = Fn.TAP, None
        out.backward(gy.cuda())
        # ---- codes: weights first (a flipped weight code contaminates every activation computed with it)
        key = f"block{bi}"
        O.EXACT_GEMM = True                 # W_q^T W_k of the QKR weight quantizer: the correctly rounded product
        try:
            wflips = _weight_flips(taps, qkr, {k: v.detach() for k, v in P.items()}, pre, wb, rep, key)
        finally:
            O.EXACT_GEMM = False
        sites, _ = _sites_of_block(taps, qkr, swin)
        dirty = mask_flips = 0
        for name, codes, v in sites:
            if name.endswith(("quan_a_qkx_fn", "quan_a_q_fn")):
                dirty += sum(n for k, n in wflips.items() if k.startswith("attn.") and not k.endswith("proj.statsq_fn"))
            elif name.endswith("fc1.input_quant_fn"):
                dirty += wflips.get("attn.proj.statsq_fn", 0)
            elif name.endswith("fc2.input_quant_fn"):
                dirty += wflips.get("mlp.fc1.statsq_fn", 0)
            o = otaps[pre + name]
            r = _compare_codes(codes, v, o["codes"], (o["x"] / o["se"]).float(), o["lo"], o["hi"])
            if r["mismatches"] or r["mask_flips"] or dirty:
                r["downstream_of_flipped_tie"] = dirty > 0
                rep[f"{key}.{name}"] = r
            if dirty == 0:
                assert r["not_ties"] == 0 and r["mask_not_ties"] == 0, (cfg, key, name, r)
            dirty += r["mismatches"]
            mask_flips += r["mask_flips"]
        dirty += wflips.get("mlp.fc2.statsq_fn", 0)
        rep["flipped_ties"] += dirty
        # ---- values and gradients
        e_out = rel_err(out.detach().cpu(), out_ref.detach())
        grads = {"d_input": (xg.grad.cpu(), xo.grad)}
        mine = dict(blk.named_parameters())
        for k, pr in bp.items():
            if pr.grad is not None:
                grads[k[len(pre):]] = (mine[k[len(pre):]].grad.detach().cpu(), pr.grad)
        gmax = max(r_.abs().max().item() for _, r_ in grads.values())
        worst, worst_name = 0.0, ""
        for name, (m_, r_) in grads.items():
            if name.endswith(ANALYTIC_ZERO):
                assert dirty or m_.abs().max().item() <= 1e-5 * gmax, (cfg, key, name)
                continue
            e = rel_err(m_, r_)
            if (m_ - r_).abs().max().item() > 1e-5 * gmax and e > worst:
                worst, worst_name = e, name
        if dirty == 0:
            assert e_out < 1e-5, (cfg, key, e_out)
        if dirty == 0 and mask_flips == 0:      # gradients: also no straight-through mask decided by a clamp-bound tie
            rep["blocks_clean"] += 1
            assert worst < GRAD_TOL, (cfg, key, worst_name, worst)
            if e_out > rep["worst_clean_block_out"]:
                rep["worst_clean_block_out"] = e_out
            if worst > rep["worst_clean_grad"]:
                rep["worst_clean_grad"], rep["worst_clean_grad_name"] = worst, f"{key}.{worst_name}"
        else:
            rep["blocks_with_flipped_tie"] += 1
            rep[f"{key}.contaminated"] = {"block_out_rel_err": e_out, "worst_grad_rel_err": worst, "worst_grad": worst_name,
                                          "flipped_codes": dirty, "flipped_masks": mask_flips}
    from ofq_b200.quantization.functional import BWD_MODE
    _report(cfg, f"teacher_forced[{BWD_MODE}]", rep)
    # ties are rare events per code, but a block of Swin-T's last stage quantizes 14 M W_qk weights: a good part of the
    # blocks must still be free of any flipped tie and therefore strictly checked
    assert rep["blocks_clean"] >= len(blocks) // 4, rep


@pytest.mark.skipif(os.environ.get("OFQ_PARITY_CHILD") == "1", reason="already the bf16x2 child run")
def test_teacher_forced_parity_with_bf16x2_backward_operands():
    """The same per-block parity run with the backward GEMM operands as two bf16 planes (OFQ_BWD_MODE=bf16x2, the
    pre-fp16 default: no range scaling, ~16 mantissa bits) instead of range-scaled fp16. The mode is read at import, so the
    run happens in a child interpreter; its per-block report lands next to the fp16 one (teacher_forced[bf16x2])."""
    import subprocess
    import sys
    env = dict(os.environ, OFQ_BWD_MODE="bf16x2", OFQ_PARITY_CHILD="1")
    r = subprocess.run([sys.executable, "-m", "pytest", str(Path(__file__)), "-x", "-q", "-m", "gpu", "-k",
                        "teacher_forced_every_block and (deit_tiny_qkr_w2a2 or deit_small_qkr_w2a2)"],
                       env=env, cwd=str(ROOT), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "2 passed" in r.stdout, r.stdout[-1000:]
