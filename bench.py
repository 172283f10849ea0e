#!/usr/bin/env python
"""Headline benchmark: QAT images/sec of the DeiT-S W2A2 attn_q (QKR) step on B200, synthetic 224x224 data.

    python bench.py --gpus N --steps K --warmup W             # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU fp32 fake-quant path (oracle port)

One "step" = forward + loss + backward (+ NCCL gradient all-reduce for N > 1) + fused AdamW update of one batch of
128 images per GPU (weak scaling: global batch 128*N; BASELINE.json config 3 is 1024 on 8 GPUs).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "QAT images/sec, DeiT-S W2A2 attn_q (QKR) step, synthetic 224x224"
UNIT = "images/s"
MODELS = {"deit_small": dict(embed_dim=384, depth=12, num_heads=6), "deit_tiny": dict(embed_dim=192, depth=12, num_heads=3),
          "swin_tiny": None}      # BASELINE.json config 4 (W3A3: --bits 3): quantized shifted-window attention, host/swin.py
# forward GFLOP per image in the quantized GEMMs of the hot path; a QAT step is 3x
GFLOP_FWD_PER_IMG = {"deit_small": 13.74, "deit_tiny": 3.00, "swin_tiny": 16.5}


BWD_DTYPE = {"f16": "range-scaled fp16", "bf16x2": "bf16 hi+lo", "bf16": "bf16"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ofq_b200", choices=["ofq_b200", "reference"])
    ap.add_argument("--model", default="deit_small", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=128, help="images per GPU")
    ap.add_argument("--bits", type=int, default=2)
    ap.add_argument("--no-qkr", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=8, help="images per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sites", default="", help="write a per-call-site kernel timing table of the instrumented steps here")
    ap.add_argument("--grad-buffer", default="none", choices=["flat", "none"],
                    help="N = 1 only (N > 1 always exchanges through the flat buffer). flat: gradients live in one flat fp32 buffer "
                         "(ofq_b200.ddp.FlatGradAllReduce); measured no faster than per-parameter gradients at N = 1 (6207 vs 6214-6302 img/s)")
    ap.add_argument("--kd", default="off", choices=["off", "fp32", "bf16"],
                    help="knowledge distillation as every reference script trains (--use-kd --kd_hard_and_soft 1, train.py:896-910): a "
                         "frozen unquantized teacher of the same architecture runs under no_grad (ofq_b200.kd.Teacher; fp32, or bf16 "
                         "autocast) and the loss is the fused KDLossSoftandHard. Not the headline workload (BASELINE's metric is the QAT step).")
    ap.add_argument("--mode", default="qat", choices=["qat", "cga", "eval"],
                    help="qat: the headline QAT step. cga: BASELINE.json config 5, the CGA fine-tune step (qk_reparam_type=1, "
                         "freeze mask fused into AdamW for every StatsQ weight, boundaryRange 0.005, lr 1e-5). eval: no-grad "
                         "inference forward (eval_scripts/deit_s/w2a2.sh uses batch 200)")
    ap.add_argument("--graph", default="on", choices=["on", "off"],
                    help="capture the whole QAT step (fwd+bwd+all-reduce+AdamW) in one CUDA graph and replay it "
                         "(ofq_b200.step_graph.CapturedStep); off = the eager Python loop the reference's train.py runs")
    ap.add_argument("--host-model", default="fused", choices=["fused", "plain"],
                    help="fused: the repo's host model (ofq_b200 LayerNorm kernels, residual add folded into the next norm). "
                         "plain: a timm-style host with torch LayerNorm and un-fused residuals (ofq_b200/host/plain.py), i.e. "
                         "the quantized modules dropped into somebody else's model, as into the reference's")
    ap.add_argument("--ddp", default="flat", choices=["flat", "bucketed"],
                    help="gradient exchange for N > 1: one flat all-reduce after backward (default: measured faster on B200, "
                         "20.87 against 21.28 ms per step at N = 2 - the persistent one-CTA-per-SM GEMMs leave the NCCL kernels "
                         "of an overlapped exchange no SM to run on, the two only take SMs from each other), or ~25 MB buckets "
                         "reduced on a communication stream while the backward is still running (torch DDP semantics, train.py:727)")
    return ap.parse_args()


def note(msg):
    if os.environ.get("OFQ_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - T0:.1f}s] {msg}", file=sys.stderr, flush=True)


T0 = time.perf_counter()


def workload_name(a):
    what = {"qat": "QAT step (fwd+bwd+AdamW)", "cga": "CGA fine-tune step (fwd+bwd+freeze-masked AdamW)",
            "eval": "eval forward (no grad)"}[getattr(a, "mode", "qat")]
    return (f"{a.model.replace('_', '-')}{'' if a.model.startswith('swin') else ' distilled'} W{a.bits}A{a.bits} attn_q {'plain' if a.no_qkr else 'QKR'} {what}, "
            f"batch {a.batch}/GPU, synthetic 224x224, random init")


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_rate(a, steps, warmup, threads):
    """The reference's CPU path for this workload, restated in oracle/ofq_oracle.py (the reference itself is
    Python that needs /root/reference and timm, neither of which exists on the GPU box): fp32 fake-quant forward,
    autograd backward and AdamW on a bounded sample of `--cpu-batch` images."""
    import torch
    import torch.nn.functional as F

    import ofq_b200.quantization as Q
    from ofq_b200.host.deit import DistilledVisionTransformer
    from oracle import ofq_oracle as O

    torch.set_num_threads(threads)
    cfg = MODELS[a.model]
    torch.manual_seed(0)
    # parameter dictionary with the reference's key names, taken from a (CPU-constructed) host model
    model = DistilledVisionTransformer(num_classes=1000, **cfg)
    names = Q.deit_qmodule_names(cfg["depth"])
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(names, a.bits, a.bits), pretrained_initialized=True,
                                             qk_reparam=not a.no_qkr)
    P = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    del model
    img = torch.randn(a.cpu_batch, 3, 224, 224)
    labels = torch.randint(0, 1000, (a.cpu_batch,))
    state = {}
    with torch.no_grad():
        O.deit_forward(img, P, cfg["depth"], cfg["num_heads"], a.bits, a.bits, not a.no_qkr, state)   # creates the LSQ scales
    params = [p for p in P.values() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=5.47e-4, weight_decay=0.05)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        cls, dist = O.deit_forward(img, P, cfg["depth"], cfg["num_heads"], a.bits, a.bits, not a.no_qkr, state)
        loss = F.cross_entropy(cls, labels) + F.cross_entropy(dist, labels)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return a.cpu_batch * len(times) / total, total / len(times)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(a.steps, 6))
    warm = max(1, min(a.warmup, 2))
    ips, sec = cpu_step_rate(a, steps, warm, threads)
    sample = f"{a.cpu_batch} images/step x {steps} steps (oracle port of the reference CPU fp32 path, torch {threads} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.index = index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # NVML directly when the bindings are there (a query takes microseconds: tens of samples inside a half-second timed
        # region); otherwise the nvidia-smi line of the profiling recipe (~0.2 s per query)
        nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            nv = (pynvml, h, bits, get_reasons)
        except Exception:
            nv = None
        while not self._stop.is_set():
            try:
                if nv is not None:
                    pynvml, h, bits, get_reasons = nv
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    r = int(get_reasons(h))
                    for n, b in bits.items():
                        if r & b:
                            self.reasons.add(n)
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append(float(out[0]))
                    self.max_mhz = float(out[1])
                    for n, v in zip(names, out[2:]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.02 if nv is not None else 0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        # the step is captured on a side stream; AccumulateGrad nodes created during the eager warm-up live on the default one
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    import ofq_b200.quantization as Q
    from ofq_b200 import _lib, ops
    from ofq_b200.cga import CGAAdamW, cga_masked_parameter_names, param_groups_weight_decay
    from ofq_b200.ddp import BucketedGradAllReduce, FlatGradAllReduce, broadcast_parameters
    from ofq_b200.host.deit import DistilledVisionTransformer
    from ofq_b200.host.plain import PlainDistilledViT
    from ofq_b200.step_graph import CapturedStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.load().ofq_device_ok() == 1, _lib.load().ofq_last_error().decode()
    note("process group up")

    from ofq_b200.quantization.functional import BWD_MODE as bwd_mode
    cfg = MODELS[a.model]
    torch.manual_seed(0)                      # identical initial weights on every rank
    swin = a.model.startswith("swin")
    if swin:
        from ofq_b200.host.swin import swin_t
        model = swin_t(num_classes=1000)
        model = Q.replace_module_by_qmodule_swin(model, Q.make_qconfigs(Q.swin_qmodule_names(), a.bits, a.bits),
                                                 pretrained_initialized=True, qk_reparam=not a.no_qkr,
                                                 qk_reparam_type=1 if a.mode == "cga" else 0).to(dev)
    else:
        model = (PlainDistilledViT if a.host_model == "plain" else DistilledVisionTransformer)(num_classes=1000, **cfg)
        names = Q.deit_qmodule_names(cfg["depth"])
        model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(names, a.bits, a.bits), pretrained_initialized=True,
                                                 qk_reparam=not a.no_qkr, qk_reparam_type=1 if a.mode == "cga" else 0).to(dev)
    def make_fp_model():          # the KD teacher: the unquantized network of the same architecture (train.py:428-442)
        if swin:
            from ofq_b200.host.swin import swin_t
            return swin_t(num_classes=1000)
        return DistilledVisionTransformer(num_classes=1000, **cfg)

    gen = torch.Generator().manual_seed(1234 + rank)
    B = a.batch
    h_img = torch.randn(B, 3, 224, 224, generator=gen).pin_memory()
    h_lbl = torch.randint(0, 1000, (B,), generator=gen).pin_memory()
    d_img, d_lbl = h_img.to(dev), h_lbl.to(dev)

    model.eval()
    with torch.no_grad():
        model(d_img)                           # setup_alpha (train.py:997-1010): creates the LSQ step sizes
    if world > 1:                              # the step sizes are data dependent: make them identical on all ranks
        broadcast_parameters(model, 0)
    model.train(a.mode != "eval")
    if a.mode == "cga":       # cga.py:953-1013: every StatsQ-quantized block weight is freeze-masked inside the AdamW kernel
        pd = dict(model.named_parameters())
        masked = [pd[n] for n in cga_masked_parameter_names(model, qk_reparam=not a.no_qkr)]
        opt = CGAAdamW(param_groups_weight_decay(model, 0.05, getattr(model, 'no_weight_decay', lambda: set())()), lr=1e-5, masked=masked,
                       wq_bitw=a.bits, boundary_range=0.005)
    else:
        opt = CGAAdamW(param_groups_weight_decay(model, 0.05, getattr(model, 'no_weight_decay', lambda: set())()), lr=5.47e-4)
    # data-parallel gradient exchange (ofq_b200/ddp.py): ONE NCCL all-reduce (mean) of a flat fp32 gradient buffer
    # (`--grad-buffer flat` uses the same flat buffer as the gradient arena at N = 1)
    ddp = None
    if a.mode != "eval" and (world > 1 or a.grad_buffer == "flat"):
        ddp = BucketedGradAllReduce(model, world) if (a.ddp == "bucketed" and world > 1) else FlatGradAllReduce(model.parameters(), world)
    flat = ddp.flat if ddp is not None else None
    teacher = kd_loss_fn = None
    if a.kd != "off" and a.mode != "eval":
        from ofq_b200.kd import Teacher
        from ofq_b200.quantization.utils import KDLossSoftandHard
        torch.manual_seed(1234)
        teacher = Teacher(make_fp_model().to(dev), dtype=torch.bfloat16 if a.kd == "bf16" else None)
        kd_loss_fn = KDLossSoftandHard()

    def step(img, lbl):
        if a.mode == "eval":
            with torch.no_grad():
                out, _ = model(img)              # eval: the mean of the two heads (deit.py:60-67)
            return out.float().mean()
        if ddp is None:
            opt.zero_grad(set_to_none=True)
        else:
            ddp.zero()
        if teacher is not None:
            out, _ = model(img)
            loss = kd_loss_fn(out, lbl, teacher(img))
        elif swin:
            logits, _ = model(img)
            loss = F.cross_entropy(logits, lbl)
        else:
            (cls, dst), _ = model(img)
            loss = F.cross_entropy(cls, lbl) + F.cross_entropy(dst, lbl)
        (ddp.scale_loss(loss) if ddp is not None else loss).backward()
        if ddp is not None:
            ddp.reduce()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    note("model built, scales initialised")
    for _ in range(a.warmup):
        step(d_img, d_lbl)
    barrier()
    note("eager warm-up done")
    run = step
    launches_per_step = None
    if a.graph == "on":
        counted = {}

        def pre_capture():
            barrier()
            if flat is None:
                opt.zero_grad(set_to_none=True)
            counted["l0"] = ops.LAUNCHES

        captured = CapturedStep(step, (d_img, d_lbl), warmup=2, pre_capture=pre_capture)
        graph, static_loss = captured.graph, captured.static_output
        static_img, static_lbl = captured.static_inputs
        launches_per_step = ops.LAUNCHES - counted["l0"]
        note("graph captured")
        barrier()

        def run(img, lbl):
            if img is not static_img:
                static_img.copy_(img, non_blocking=True)
                static_lbl.copy_(lbl, non_blocking=True)
            graph.replay()
            return static_loss

        d_img, d_lbl = static_img, static_lbl
        for _ in range(2):
            run(d_img, d_lbl)
    note("timed region starts")
    # ---- timed region 1: inputs resident in HBM
    barrier()
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        t_host0 = time.perf_counter()
        for _ in range(a.steps):
            run(d_img, d_lbl)
        host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / a.steps
        e1.record()
        barrier()
    launches = ops.LAUNCHES - launches0 if launches_per_step is None else launches_per_step * a.steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    # ---- timed region 2: end to end (pinned host -> device every step, loss read back every step)
    barrier()
    e0.record()
    last = 0.0
    if a.graph == "on":
        # input pipeline of a training loop: the pinned-host -> device copy of step i+1's batch runs on a copy stream while
        # step i computes; every step still copies its own inputs inside the timed region and reads its loss back
        cur = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        stage_img, stage_lbl = torch.empty_like(static_img), torch.empty_like(static_lbl)

        def prefetch(after):
            copy_stream.wait_event(after)               # the staging buffers are free again
            with torch.cuda.stream(copy_stream):
                stage_img.copy_(h_img, non_blocking=True)
                stage_lbl.copy_(h_lbl, non_blocking=True)
                return copy_stream.record_event()

        ready = prefetch(cur.record_event())
        for i in range(a.steps):
            cur.wait_event(ready)
            static_img.copy_(stage_img, non_blocking=True)
            static_lbl.copy_(stage_lbl, non_blocking=True)
            consumed = cur.record_event()
            graph.replay()
            if i + 1 < a.steps:
                ready = prefetch(consumed)
            last = static_loss.item()
    else:
        for _ in range(a.steps):
            last = run(h_img.to(dev, non_blocking=True), h_lbl.to(dev, non_blocking=True)).item()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    # ---- instrumented steps: per-kernel CUDA-event timing on the launching stream (share of the step per family)
    roof = None
    nprof = 3
    if rank == 0:
        ops.PROFILE = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(nprof):              # eager (un-graphed) steps on every rank; only rank 0 records events
        step(d_img, d_lbl)
    ev1.record()
    barrier()
    if rank == 0:
        fam, sites = {}, {}
        for name, s0, s1, nbytes, nflops, tag in ops.PROFILE:
            dt = s0.elapsed_time(s1)
            for d, k in ((fam, name), (sites, f"{name} {tag}".strip())):
                f = d.setdefault(k, [0.0, 0, 0.0, 0.0])
                f[0] += dt
                f[1] += 1
                f[2] += nbytes
                f[3] += nflops
        ops.PROFILE = None
        if a.sites:
            with open(a.sites, "w") as fh:
                for k, v in sorted(sites.items(), key=lambda kv: -kv[1][0]):
                    fh.write(f"{v[0] / nprof:9.4f} ms/step  {v[1] / nprof:6.1f} launches  {v[0] / v[1] * 1e3:8.1f} us  "
                             f"{v[2] / (v[0] * 1e-3) / 1e9:8.1f} GB/s  {v[3] / (v[0] * 1e-3) / 1e12:8.1f} TFLOP/s  {k}\n")
        step_ms = ev0.elapsed_time(ev1) / nprof
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            j = json.loads(pk.read_text())
            peaks = {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                     "source": "MEASURED_PEAKS.json (copy GB/s; sustained cuBLAS bf16)"}
        top = max(fam.items(), key=lambda kv: kv[1][0])
        name, (tms, cnt, nbytes, nflops) = top
        t_hbm = nbytes / (peaks["hbm_gbs"] * 1e9)
        tensor_peak = peaks["bf16_tflops"] * (2.0 if name == "gemm_i8" else 1.0) * 1e12     # int8 = 2x the bf16 rate
        t_tensor = nflops / tensor_peak
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            traffic = json.loads(tf.read_text()).get(name)
        if t_hbm >= t_tensor:
            ach = nbytes / (tms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
        else:
            ach = nflops / (tms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": tensor_peak / 1e12, "unit": "TFLOP/s", "frac": ach / (tensor_peak / 1e12)}
        roof.update({"traffic": traffic, "kernel": name, "launches_per_step": cnt / nprof, "avg_launch_us": tms / cnt * 1e3,
                     "share_of_step": (tms / nprof) / (ms_total / a.steps), "peak_source": peaks["source"],
                     "families_ms_per_step": {k: round(v[0] / nprof, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])},
                     "families_gbps": {k: round(v[2] / (v[0] * 1e-3) / 1e9, 1) for k, v in fam.items() if v[0] > 0},
                     "instrumented_step_ms": step_ms})
    def finish():
        # NCCL communicators that were captured into a CUDA graph do not tear down reliably: synchronise, agree that
        # every rank is done, then leave without running destructors (the JSON line is already flushed)
        if world > 1:
            barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        return finish()
    cpu = None
    if world == 1 and not a.no_cpu_baseline and not swin:
        threads = os.cpu_count() or 1
        ips, sec = cpu_step_rate(a, 4, 1, threads)
        cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{a.cpu_batch} images/step x 4 steps of the same model (oracle port, torch {threads} threads, {sec:.2f} s/step)"}
    imgs = B * world * a.steps
    value = imgs / (ms_total * 1e-3)
    e2e = imgs / (ms2.item() * 1e-3)
    flops_step = (1 if a.mode == "eval" else 3) * GFLOP_FWD_PER_IMG[a.model] * 1e9 * B
    line = {
        "metric": METRIC if a.mode == "qat" else METRIC.replace("QAT images/sec", f"{a.mode} images/sec").replace(" step,", f" {a.mode},"),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": f"int8 codes fwd (s32 accumulate) / {BWD_DTYPE[bwd_mode]} bwd (f32 accumulate), f32 activations",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush",
                   "optimizer": {"qat": "fused AdamW lr 5.47e-4 wd 0.05", "cga": "fused CGA-masked AdamW lr 1e-5 wd 0.05 BR 0.005", "eval": "none"}[a.mode], "cuda_graph": a.graph == "on", "host_model": a.host_model, "ddp": (a.ddp if world > 1 else None), "grad_buffer": a.grad_buffer, "kd": a.kd, "host_enqueue_ms_per_step": host_enqueue_ms, "bwd_mode": bwd_mode,
                   "quantized_gemm_tflops_per_gpu": flops_step / (ms_total / a.steps * 1e-3) / 1e12},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": (h_img.numel() * 4 + h_lbl.numel() * 8) * world,
                "d2h_bytes_per_step": 4 * world, "last_loss": last},
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
