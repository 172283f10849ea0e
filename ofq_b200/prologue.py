"""Step prologue: the weight-side work of a whole quantized model in three launches per forward pass.

StatsQ codes (statsq.py:133-150), the W_qk products of the query-key reparameterisation (attention.py:190-194) and the
effective LSQ step sizes (lsq.py:593) depend only on parameters, which change at the optimizer step and nowhere else.
The reference (and a per-layer port) recomputes each of them inside every layer's forward: 157 small launches per
DeiT-S step.  Here the layers *register* that work the first time they run under a prologue (ops.statsq_codes /
ops.wqk_compose / ops.lsq_effective_scale consult `ACTIVE`); from the next forward on, `begin()` - a forward pre-hook
of the root model installed by replace_module_by_qmodule_{deit,swin} - produces all of it with the multi-tensor
kernels (ofq_wqk_compose_multi, ofq_statsq_codes_multi, ofq_lsq_effective_scale_multi) into persistent buffers and
the layers pick their results up.  Layers called outside a root forward, or with operands that were not registered
(new batch size -> new gradient-scale factor, re-allocated parameter), compute as before.

The persistent buffers are what autograd saves for the backward; they are rewritten by the next forward, which
yields the same values unless the optimizer stepped in between (and then the old graph is stale anyway).
"""
from __future__ import annotations

import os
import struct
from typing import Optional

import torch

from . import _lib

ACTIVE: Optional["StepPrologue"] = None
ENABLED = os.environ.get("OFQ_PROLOGUE", "1") != "0"          # tests / A-B measurements switch the whole mechanism off here

# Bumped by optimizers that write parameters through raw pointers (CGAAdamW: torch's per-tensor version counters do not see
# those writes). Together with the version counters it forms the "weights changed" signature of a prologue.
WEIGHT_EPOCH = 0


def weights_changed() -> None:
    """Tell every prologue that parameters were modified behind autograd's back (raw-pointer kernels, `p.data` tricks)."""
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1


def current_generation():
    """(prologue, generation) when the running forward is being served from persistent prologue buffers, else None: what an
    autograd node stores to detect, in its backward, that a later forward re-produced those buffers for OTHER weights."""
    p = ACTIVE
    return (p, p.generation) if (p is not None and p.fresh) else None


def check_generation(token) -> None:
    if token is not None and token[0].generation != token[1]:
        raise RuntimeError("ofq_b200: this graph saved weight codes / step sizes that live in the step prologue's persistent buffers, "
                           "and a later forward re-produced them after the weights changed; run backward before the next "
                           "optimizer step + forward, or disable the prologue (OFQ_PROLOGUE=0) for such graphs")


class _Job:
    __slots__ = ("key", "tensors", "out", "used", "meta")

    def __init__(self, key, tensors, out, meta=None):
        self.key, self.tensors, self.out, self.used, self.meta = key, tensors, out, True, meta


class StepPrologue:
    def __init__(self):
        self.fresh = False
        self.scale, self.statsq, self.wqk = {}, {}, {}
        self.dirty = True
        self.tables = None
        self.launches = 0
        self.frozen = False        # export.load_packed: the buffers hold deployed codes, the weight-side kernels never run
        self.generation = 0        # bumped whenever the buffers are re-produced for changed weights
        self._sig = None           # weights signature of the last production
        self.cache_hits = 0

    def _signature(self):
        v = 0
        for d in (self.scale, self.statsq, self.wqk):
            for j in d.values():
                for t in j.tensors:
                    if t is not None:
                        v += t._version
        return (WEIGHT_EPOCH, v)

    def __deepcopy__(self, memo):
        # copy.deepcopy(model) (e.g. timm's ModelEmaV2): the copy gets a prologue of its own; jobs hold raw pointers of
        # the original's parameters and must not travel
        return StepPrologue()

    # ------------------------------------------------------------------ registration / lookup (called from ops)
    def get_scale(self, alpha: torch.Tensor, g: float, recip: bool):
        key = (alpha.data_ptr(), alpha.numel(), float(g), bool(recip))
        job = self.scale.get(key)
        if job is None:
            out = torch.empty(((2,) if recip else ()) + tuple(alpha.shape), dtype=torch.float32, device=alpha.device)
            job = self.scale[key] = _Job(key, (alpha,), out)
            self.dirty = True
            return job, False
        job.used = True
        return job, self.fresh

    def get_statsq(self, w, bits, aft, bias, fmt16=None):
        R, Cc = w.shape
        key = (w.data_ptr(), R, Cc, int(bits), 0 if aft is None else aft.data_ptr(), 0 if bias is None else bias.data_ptr(),
               fmt16)
        job = self.statsq.get(key)
        if job is None:
            dev = w.device
            out = dict(codes=torch.empty((R, Cc), dtype=torch.int8, device=dev),
                       cs2=torch.empty((2, R), dtype=torch.float32, device=dev),
                       colterm=torch.empty(R, dtype=torch.float32, device=dev) if (aft is not None or bias is not None) else None,
                       codes16=None if fmt16 is None else torch.empty((R, Cc), device=dev,
                                                                      dtype=torch.float16 if fmt16 == 1 else torch.bfloat16))
            job = self.statsq[key] = _Job(key, (w, aft, bias), out, meta=(int(bits), fmt16))
            self.dirty = True
            return job, False
        job.used = True
        return job, self.fresh

    def get_wqk(self, wq, wk, H):
        key = (wq.data_ptr(), wk.data_ptr(), int(H), wq.shape[0], wq.shape[1])
        job = self.wqk.get(key)
        if job is None:
            Cc = wq.shape[1]
            out = torch.empty((H * Cc, Cc), dtype=torch.float32, device=wq.device)
            job = self.wqk[key] = _Job(key, (wq, wk), out, meta=int(H))
            self.dirty = True
            return job, False
        job.used = True
        return job, self.fresh

    # ------------------------------------------------------------------ per-forward
    def _upload(self, blob: bytes, dev):
        host = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
        t = host.to(dev, non_blocking=True)
        t._ofq_host = host
        return t

    def _build(self):
        dev = None
        for d in (self.scale, self.statsq, self.wqk):
            for j in d.values():
                dev = j.tensors[0].device
                break
            if dev is not None:
                break
        tb = {}
        if self.wqk:
            groups = {}
            for j in self.wqk.values():
                wq = j.tensors[0]
                groups.setdefault((j.meta, wq.shape[0] // j.meta, wq.shape[1]), []).append(j)
            tb["wqk"] = []
            for (H, hd, Cc), jobs in groups.items():
                blob = b"".join(struct.pack("<QQQ", j.tensors[0].data_ptr(), j.tensors[1].data_ptr(), j.out.data_ptr()) for j in jobs)
                tb["wqk"].append((self._upload(blob, dev), len(jobs), H, hd, Cc))
        if self.statsq:
            blob, first = b"", 0
            for j in self.statsq.values():
                w, aft, bias = j.tensors
                R, Cc = w.shape
                o = j.out
                bits, fmt16 = j.meta
                blob += struct.pack("<QQQQQQQQqiifiii", w.data_ptr(), 0 if aft is None else aft.data_ptr(),
                                    0 if bias is None else bias.data_ptr(), o["codes"].data_ptr(), o["cs2"][0].data_ptr(),
                                    o["cs2"][1].data_ptr(), 0 if o["colterm"] is None else o["colterm"].data_ptr(),
                                    0 if o["codes16"] is None else o["codes16"].data_ptr(),
                                    w.stride(0), R, Cc, float(1 << (bits - 1)), first, 1 if fmt16 == 1 else 0, 0)
                first += (R + 7) // 8
            tb["statsq"] = (self._upload(blob, dev), len(self.statsq), first)
        if self.scale:
            blob, first = b"", 0
            for j in self.scale.values():
                a = j.tensors[0]
                recip = j.key[3]
                out0 = j.out[0] if recip else j.out
                blob += struct.pack("<QQQifii", a.data_ptr(), out0.data_ptr(), j.out[1].data_ptr() if recip else 0, a.numel(),
                                    j.key[2], first, 0)
                first += (a.numel() + 255) // 256
            tb["scale"] = (self._upload(blob, dev), len(self.scale), first)
        self.tables = tb
        self.dirty = False

    def begin(self, device=None):
        """Root forward pre-hook: run the registered work (if any) and make it available to the layers."""
        global ACTIVE
        ACTIVE = self
        self.fresh = False
        from . import ops
        ops.arena_begin(torch.cuda.current_device() if device is None else device)
        if not ENABLED or not (self.scale or self.statsq or self.wqk):
            return
        for d in (self.scale, self.statsq, self.wqk):       # a job must still describe live storage of the same layout
            for j in d.values():
                j.used = False
        rebuilt = self.dirty
        if self.dirty:
            self._build()
        sig = self._signature()
        if sig != self._sig:
            self.generation += 1
        elif not rebuilt and not torch.is_grad_enabled():
            # inference with unchanged weights (same optimizer epoch, same version counters, same registered jobs): the
            # persistent buffers still hold this model's codes / W_qk / step sizes - nothing to launch. (Weights modified
            # through `p.data` or foreign raw-pointer kernels are invisible to both counters: call weights_changed().)
            self.cache_hits += 1
            self.fresh = True
            return
        self._sig = sig
        lib = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        tb = self.tables
        for table, n, H, hd, Cc in (() if self.frozen else tb.get("wqk", ())):
            ops._call("wqk_compose", 1, 4.0 * n * (2 * H * hd * Cc + H * Cc * Cc), 2.0 * n * H * hd * Cc * Cc, lib.ofq_wqk_compose_multi,
                      table.data_ptr(), n, H, hd, Cc, st)
        if "statsq" in tb and not self.frozen:
            table, n, blocks = tb["statsq"]
            nbytes = sum(5.0 * j.tensors[0].numel() for j in self.statsq.values())
            ops._call("statsq", 1, nbytes, 0, lib.ofq_statsq_codes_multi, table.data_ptr(), n, blocks, st)
        if "scale" in tb:
            table, n, blocks = tb["scale"]
            ops._call("lsq_scale", 1, 12.0 * sum(j.tensors[0].numel() for j in self.scale.values()), 0,
                      lib.ofq_lsq_effective_scale_multi, table.data_ptr(), n, blocks, st)
        self.fresh = True

    def end(self):
        """Root forward hook: results are only valid inside the forward that produced them; drop jobs nobody asked for."""
        global ACTIVE
        if self.fresh:
            for d in (self.scale, self.statsq, self.wqk):
                dead = [k for k, j in d.items() if not j.used]
                for k in dead:
                    del d[k]
                    self.dirty = True
        self.fresh = False
        ACTIVE = None


def install(model: torch.nn.Module) -> StepPrologue:
    """Attach a StepPrologue to `model` (idempotent): begin() before its forward, end() after."""
    pro = getattr(model, "_ofq_prologue", None)
    if pro is not None:
        return pro
    pro = StepPrologue()
    model._ofq_prologue = pro
    # the hooks look the prologue up on the module they fire for: a deep copy of the model runs its own
    model.register_forward_pre_hook(_begin_hook)
    model.register_forward_hook(_end_hook, always_call=True)
    return pro


def _begin_hook(m, inp):
    pro = getattr(m, "_ofq_prologue", None)
    if pro is not None and inp and torch.is_tensor(inp[0]) and inp[0].is_cuda:
        pro.begin(inp[0].device)


def _end_hook(m, inp, out):
    pro = getattr(m, "_ofq_prologue", None)
    if pro is not None:
        pro.end()
