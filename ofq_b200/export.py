"""Deployment export (the reference has no counterpart - it stops at fake-quant floats).

`export_packed(model, example)` runs one inference forward and collects, for every StatsQ-quantized weight of the model
(QLinear weights, the V projection and the composite W_q^T W_k products of the query-key reparameterisation - the products
themselves, so the deployed model needs neither W_q / W_k nor the compose kernel), the bit-exact integer codes packed at their
true width (2 / 3 / 4 bits per weight, `ofq_pack_codes`), the per-output-channel scale and the folded shift / bias term.
Everything else the forward needs (LSQ step sizes, shifts, LayerNorm, embeddings, the 8-bit ends) is kept as ordinary tensors.
DeiT-S W2A2: 21.2 M quantized weights -> 5.3 MB of codes instead of 85 MB of fp32.

`load_packed(model, packed, example)` installs the unpacked codes into the step prologue's persistent buffers of a freshly
built model and FREEZES the prologue: the integer inference forward (int8 tensor-core GEMMs, fused attention) then runs from the
packed codes alone - the fp32 weights are never read again (`drop_fp32=True` zeroes them to prove it).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib, ops

FORMAT = "ofq_b200-packed-v1"


def _st():
    return torch.cuda.current_stream().cuda_stream


def pack_codes(codes: torch.Tensor, bits: int) -> torch.Tensor:
    """int8 StatsQ codes [R, C] (odd integers 2k+1) -> uint8 [R, ceil(C / 8) * bits]."""
    assert codes.is_cuda and codes.dtype == torch.int8 and codes.dim() == 2 and codes.stride(1) == 1
    lib = _lib.load()
    R, Cc = codes.shape
    out = torch.empty((R, lib.ofq_packed_row_bytes(Cc, bits)), dtype=torch.uint8, device=codes.device)
    ops._call("pack_codes", 1, R * Cc * (1.0 + bits / 8.0), 0, lib.ofq_pack_codes, codes.data_ptr(), R, Cc, codes.stride(0), bits,
              out.data_ptr(), out.stride(0), _st())
    return out


def unpack_codes(packed: torch.Tensor, cols: int, bits: int, out: torch.Tensor = None) -> torch.Tensor:
    assert packed.is_cuda and packed.dtype == torch.uint8 and packed.dim() == 2 and packed.stride(1) == 1
    R = packed.shape[0]
    if out is None:
        out = torch.empty((R, cols), dtype=torch.int8, device=packed.device)
    ops._call("unpack_codes", 1, R * cols * (1.0 + bits / 8.0), 0, _lib.load().ofq_unpack_codes, packed.data_ptr(), packed.stride(0), R,
              cols, bits, out.data_ptr(), out.stride(0), _st())
    return out


def _sites(model: torch.nn.Module):
    """[(site name, statsq job, [names of the fp32 parameters the site replaces])] of the model's (registered) prologue."""
    pro = getattr(model, "_ofq_prologue", None)
    if pro is None or not pro.statsq:
        raise RuntimeError("export needs a model wrapped by replace_module_by_qmodule_* that has run an inference forward")
    by_ptr = {p.data_ptr(): n for n, p in model.named_parameters()}
    wqk_out = {}
    for j in pro.wqk.values():
        wq, wk = j.tensors
        nq, nk = by_ptr[wq.data_ptr()], by_ptr[wk.data_ptr()]
        wqk_out[j.out.data_ptr()] = (nq.rsplit(".", 2)[0] + ".W_qk", [nq, nk])
    sites = []
    for j in pro.statsq.values():
        w = j.tensors[0]
        if w.data_ptr() in by_ptr:
            name = by_ptr[w.data_ptr()]
            sites.append((name, j, [name]))
        elif w.data_ptr() in wqk_out:
            name, repl = wqk_out[w.data_ptr()]
            sites.append((name, j, repl))
        else:
            raise RuntimeError("a StatsQ job of the prologue does not belong to a parameter of this model")
    return pro, sites


@torch.no_grad()
def export_packed(model: torch.nn.Module, example: torch.Tensor) -> Dict:
    model.eval()
    model(example)
    model(example)                      # second forward: every job is registered and served by the prologue
    pro, sites = _sites(model)
    out = {"format": FORMAT, "sites": {}, "state": {}}
    replaced = set()
    for name, j, repl in sites:
        if j.meta[1] is not None:
            continue                    # training-mode twin of a job (fp16 copy requested): same codes
        bits = j.meta[0]
        codes = j.out["codes"]
        out["sites"][name] = {"bits": bits, "rows": codes.shape[0], "cols": codes.shape[1], "packed": pack_codes(codes, bits).cpu(),
                              "colscale": j.out["cs2"][0].clone().cpu(),
                              "colterm": None if j.out["colterm"] is None else j.out["colterm"].clone().cpu()}
        replaced.update(repl)
    for k, v in model.state_dict().items():
        if k not in replaced:
            out["state"][k] = v.detach().clone().cpu()
    out["replaced"] = sorted(replaced)
    return out


def packed_nbytes(packed: Dict) -> Dict[str, int]:
    codes = sum(s["packed"].numel() for s in packed["sites"].values())
    scales = sum(4 * (s["colscale"].numel() + (0 if s["colterm"] is None else s["colterm"].numel())) for s in packed["sites"].values())
    rest = sum(v.numel() * v.element_size() for v in packed["state"].values())
    return {"codes": codes, "scales": scales, "other_state": rest, "quantized_weights": sum(s["rows"] * s["cols"] for s in packed["sites"].values())}


@torch.no_grad()
def load_packed(model: torch.nn.Module, packed: Dict, example: torch.Tensor, drop_fp32: bool = True) -> torch.nn.Module:
    """`model`: built like the exported one (same architecture, replace_module_by_qmodule_* applied), on the GPU."""
    assert packed.get("format") == FORMAT, "not an ofq_b200 packed export"
    dev = example.device
    # (on the model's device: the lazily created LSQ step sizes are adopted as they come, lsq.py:541)
    missing, unexpected = model.load_state_dict({k: v.to(dev) for k, v in packed["state"].items()}, strict=False)
    assert set(missing) <= set(packed["replaced"]) and not unexpected, (missing, unexpected)
    model.eval()
    model(example)
    model(example)                      # registers the prologue jobs (their buffers are what the layers read from now on)
    pro, sites = _sites(model)
    names = {n for n, _, _ in sites}
    assert set(packed["sites"]) <= names, sorted(set(packed["sites"]) - names)
    params = dict(model.named_parameters())
    for name, j, repl in sites:
        if name not in packed["sites"]:
            continue
        s = packed["sites"][name]
        assert j.meta[0] == s["bits"] and tuple(j.out["codes"].shape) == (s["rows"], s["cols"]), name
        unpack_codes(s["packed"].to(dev), s["cols"], s["bits"], out=j.out["codes"])
        cs = s["colscale"].to(dev)
        j.out["cs2"][0].copy_(cs)
        j.out["cs2"][1].copy_(1.0 / cs)
        if s["colterm"] is not None:
            j.out["colterm"].copy_(s["colterm"].to(dev))
        if drop_fp32:
            for n in repl:
                params[n].zero_()
    pro.frozen = True                   # weight-side kernels never run again: the codes above ARE the weights
    return model
