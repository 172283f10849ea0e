"""GPU parity at the REAL width and depth of every BASELINE.json configuration (depth-12 DeiT-T / DeiT-S, Swin-T stage
dimensions; batch 8), against (a) golden outputs of the unmodified reference (tests/golden/full_*.npz) and (b) the CPU
oracle run live on the same regenerated parameters (the oracle is pinned to those goldens by test_oracle_fullsize.py).

* free running (whole model, reference goldens): logits, loss, eval logits <= 1e-3; every block output (sampled) reported
  per block so that depth compounding is visible; every gradient (sampled) <= 1e-3 relative (norm-wise, with the absolute
  escape of test_gpu_layers.check_grads for gradients that are round-off in the reference); integer codes of the first and
  last block against the reference's: mismatch counts are reported (a flipped rounding tie upstream changes a whole row of
  a 2-bit layer downstream, so free-running codes of block 11 are informative, not a pass criterion).
* teacher forced (first and last block on the reference's own block input): every integer code tensor of the block
  (activation codes of all quantizers, StatsQ weight codes) must equal the oracle's bit for bit; a mismatch is accepted
  only as a PROVEN TIE: the two pre-round values differ by no more than 2e-5 * max(1, |v|) (fp32 GEMM round-off of the
  reference's sgemm against the exact integer GEMM) and therefore straddle a rounding boundary. Sites downstream of a
  flipped tie inside the same module call are reported, not asserted.

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
TIE_TOL = 2e-5
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


def _compare_codes(mine, v_mine, ref, v_ref=None):
    """Mismatch statistics of one code tensor. With pre-round values of both sides: which mismatches are proven ties."""
    mine = mine.detach().cpu().to(torch.int16).reshape(ref.shape)
    bad = mine != ref.to(torch.int16)
    n = int(bad.sum())
    r = {"numel": ref.numel(), "mismatches": n}
    if v_ref is not None:
        vm = v_mine.detach().cpu().reshape(ref.shape).float()
        d = (vm - v_ref).abs()
        inside = v_ref.abs() < 1e3
        r["max_pre_round_diff"] = float((d[inside] / v_ref[inside].abs().clamp_min(1.0)).max()) if bool(inside.any()) else 0.0
        if n:
            tol = TIE_TOL * v_ref[bad].abs().clamp_min(1.0)
            ties = d[bad] <= tol
            r["ties"] = int(ties.sum())
            r["not_ties"] = n - int(ties.sum())
        else:
            r["ties"] = r["not_ties"] = 0
    return r


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
    rep["grad_failures"] = failures
    model.eval()
    with torch.no_grad():
        ev = model(img)[0]
    rep["eval_logits_rel_err"] = rel_err(ev.cpu(), g["eval_logits"])
    from ofq_b200.quantization.functional import BWD_MODE
    _report(cfg, f"free_running[{BWD_MODE}]", rep)
    assert max(rep["logits_rel_err"]) < OUT_TOL, rep["logits_rel_err"]
    assert abs(rep["loss"] - rep["loss_ref"]) <= OUT_TOL * abs(rep["loss_ref"])
    assert rep["eval_logits_rel_err"] < OUT_TOL
    assert max(rep["block_out_rel_err"]) < OUT_TOL, rep["block_out_rel_err"]
    if nimg > 0:      # first block: its inputs differ from the reference's by fp32 round-off only
        first_bad = sum(v["mismatches"] for k, v in rep["codes_vs_reference"].items() if k.startswith("first."))
        first_all = sum(v["numel"] for k, v in rep["codes_vs_reference"].items() if k.startswith("first."))
        assert first_bad <= 1e-4 * first_all, (first_bad, first_all)
    assert not failures, failures[:8]


def _swin_tap_start(bi, qkr):
    """Index of the first tap of Swin block `bi` (12 blocks over stages of depth 2, 2, 6, 2; one `reduction` QLinear tap
    after each of the first three stages)."""
    per = 4 if qkr else 5
    stage_end = (2, 4, 10, 12)
    reductions = sum(1 for e in stage_end[:3] if bi >= e)
    return bi * per + reductions


# ------------------------------------------------------------------------------------------------ teacher forced
def _oracle_block(cfg, P, pre, x, bi):
    from oracle import ofq_oracle as O
    model_name, wb, ab, qkr, _, _ = FC.CONFIGS[cfg]
    if model_name != "swin_tiny":
        return O.deit_block(x, P, pre, FC.MODEL_DIMS[model_name]["num_heads"], wb, ab, qkr)
    stage = next(i for i, e in enumerate((2, 4, 10, 12)) if bi < e)
    j = bi - (0, 2, 4, 10)[stage]
    heads = (3, 6, 12, 24)[stage]
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), P[pre + "norm1.weight"], P[pre + "norm1.bias"], 1e-5)
    shift = (0, 0) if j % 2 == 0 else (3, 3)
    x = x + O.swin_window_attention(h, P, pre + "attn.", heads, wb, ab, qkr, (7, 7), shift)
    h = F.layer_norm(x, (C,), P[pre + "norm2.weight"], P[pre + "norm2.bias"], 1e-5)
    return x + O.qmlp(h, P, pre + "mlp.", wb, ab)


@pytest.mark.parametrize("cfg", [c for c in FC.CONFIGS if FC.CONFIGS[c][5] > 0])
def test_teacher_forced_codes_against_oracle(Fn, cfg):
    from oracle import ofq_oracle as O
    g = load_golden(f"full_{cfg}")
    model_name, wb, ab, qkr, _, nimg = FC.CONFIGS[cfg]
    swin = model_name == "swin_tiny"
    model = FC.load_repo_model(cfg, g).cuda().train()
    P = FC.oracle_params(cfg, g, requires_grad=False)
    blocks = _block_modules(model, model_name)
    prefixes = FC.block_prefixes(model_name)
    rep, hard_fail = {}, []
    for tag, bi in (("first", 0), ("last", len(blocks) - 1)):
        x_in = g[f"{tag}.block_in"]
        O.TAPS = {}
        try:
            with torch.no_grad():
                out_ref = _oracle_block(cfg, P, prefixes[bi], x_in, bi)
            otaps = O.TAPS
        finally:
            O.TAPS = None
        blk = blocks[bi]
        Fn.TAP = []
        try:
            xg = x_in.cuda().requires_grad_(True)          # grad mode: the probabilities are kept (they are the pre-round value)
            if swin:
                out = blk(xg)
            else:
                blk._defer = False
                out, _ = blk(xg)
        finally:
            taps, Fn.TAP = Fn.TAP, None
        rep[f"{tag}.block_out_rel_err"] = rel_err(out.detach().cpu(), out_ref)
        sites, _ = _sites_of_block(taps, qkr, swin)
        upstream_flips = 0
        for name, codes, v in sites:
            o = otaps[prefixes[bi] + name]
            v_ref = (o["x"] / o["se"]).float()
            r = _compare_codes(codes, v, o["codes"], v_ref)
            r["downstream_of_flipped_tie"] = upstream_flips > 0
            rep[f"{tag}.{name}"] = r
            if r["not_ties"] and upstream_flips == 0:
                hard_fail.append((tag, name, r))
            if name.endswith("fc1.input_quant_fn"):
                upstream_flips = 0                           # the MLP branch starts from the block's own residual stream ...
            upstream_flips += r["mismatches"]                # ... everything after a flipped code inside a branch is contaminated
        for name, wc in _weight_sites(taps, qkr):
            if name.endswith("qk_quant"):
                a = prefixes[bi] + "attn."
                H = blocks[bi].attn.num_heads
                w = O.wqk_compose(P[a + "q.weight"], P[a + "k.weight"], H)
            elif name.endswith("v_quant"):
                w = P[prefixes[bi] + "attn.v.weight"]
            else:
                w = P[prefixes[bi] + name[: -len("statsq_fn")] + "weight"]
            ref_codes, _ = O.statsq_codes(w, wb)
            b4, _ = O.statsq_pre_round(w, wb)
            r = _compare_codes(wc, None, ref_codes)
            if r["mismatches"]:
                bad = wc.detach().cpu().to(torch.int32) != ref_codes
                frac = (b4[bad] - torch.floor(b4[bad]) - 0.5).abs()        # distance of the pre-round value from the rounding tie
                r["ties"] = int((frac <= TIE_TOL).sum())
                r["not_ties"] = r["mismatches"] - r["ties"]
                if r["not_ties"]:
                    hard_fail.append((tag, name, r))
            rep[f"{tag}.w.{name}"] = r
    _report(cfg, "teacher_forced", rep)
    assert not hard_fail, hard_fail
    assert rep["first.block_out_rel_err"] < 1e-4 and rep["last.block_out_rel_err"] < 1e-4, rep
