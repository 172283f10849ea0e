"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_* --csv` launch list per kernel family.
usage: python tools/summarize_launches.py gpurun_out/r01_launches.csv profiles/r01_launches_summary.md [profiles/traffic.json]"""
import collections, csv, json, re, sys

one_step = "--one-step" in sys.argv
if one_step:
    sys.argv.remove("--one-step")
rng = None          # --range a:b keeps launches a <= index < b of the capture (in launch order)
for a in list(sys.argv):
    if a.startswith("--range="):
        rng = tuple(int(v) for v in a[8:].split(":"))
        sys.argv.remove(a)
src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
cur = {}
for r in csv.DictReader(lines):
    scale = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
    cur.setdefault((r["ID"], r["Kernel Name"]), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * scale

def family(k):
    m = re.search(r"gemm_tc_kernel<(\d), (\d+), (\d+)(?:, (\d+), (\d+))?>", k)
    if m:
        return f"ofq gemm_tc_kernel<{'i8' if m.group(1) == '0' else '16-bit'}, BN={m.group(2)}{', dual-A' if m.group(4) == '2' else ''}>"
    m = re.search(r"gemm_tc_pair_kernel<(\d), (\d+)", k)
    if m:
        return f"ofq gemm_tc_pair_kernel<{'i8' if m.group(1) == '0' else '16-bit'}, BN={m.group(2)}>"
    m = re.search(r"ofq::(?:attn::)?(\w+)", k)
    if m:
        return "ofq " + m.group(1)
    m = re.search(r"attn::(\w+_kernel)", k)
    if m:
        return "ofq " + m.group(1)
    m = re.search(r"<unnamed>::(\w+)", k)
    if m and "at::" not in k:
        return "ofq " + m.group(1)
    m = re.search(r"at::(?:native::)?(?:<unnamed>::)?(\w+)<[^>]*?(\w+Functor|\w+KernelImpl|\w+_kernel|FillFunctor)", k)
    if m:
        return f"torch {m.group(1)}[{m.group(2)}]"
    m = re.search(r"(?:void )?([\w:]+)", k)
    return "torch " + (m.group(1) if m else k)[:50]

if one_step:   # keep exactly one QAT step: the launches between the last AdamW launch of one step and of the next
    keys = sorted(cur, key=lambda ik: int(ik[0]))
    opt = [i for i, ik in enumerate(keys) if "adamw" in ik[1]]
    ends = [i for j, i in enumerate(opt) if j + 1 == len(opt) or opt[j + 1] - i > 50]
    assert len(ends) >= 2, "the capture does not hold a whole step"
    cur = {ik: cur[ik] for ik in keys[ends[-2] + 1: ends[-1] + 1]}
if rng is not None:
    keys = sorted(cur, key=lambda ik: int(ik[0]))
    cur = {ik: cur[ik] for ik in keys[rng[0]:rng[1]]}
per = collections.defaultdict(lambda: dict(n=0, ns=0.0, rd=0.0, wr=0.0))
for (_, k), d in cur.items():
    p = per[family(k)]
    p["n"] += 1
    p["ns"] += d.get("gpu__time_duration.sum", 0)
    p["rd"] += d.get("dram__bytes_read.sum", 0)
    p["wr"] += d.get("dram__bytes_write.sum", 0)
tot = sum(p["ns"] for p in per.values())
ofq = sum(p["ns"] for k, p in per.items() if k.startswith("ofq"))
rows = sorted(per.items(), key=lambda kv: -kv[1]["ns"])
with open(dst, "w") as f:
    f.write(f"# ncu launch list summary ({src})\n\n")
    f.write(f"{len(cur)} launches, {tot / 1e6:.2f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES, not absolutes).\n")
    f.write(f"ofq_b200 kernels: {100 * ofq / tot:.1f} % of the captured kernel time; the rest is torch glue (LayerNorm, GELU, residual adds, fills, heads, patch embed).\n\n")
    f.write("| kernel family | launches | total ms | share % | avg us | DRAM MB / launch (read+write) |\n|---|---:|---:|---:|---:|---:|\n")
    for k, p in rows[:60]:
        f.write(f"| {k} | {p['n']} | {p['ns'] / 1e6:.3f} | {100 * p['ns'] / tot:.2f} | {p['ns'] / p['n'] / 1e3:.1f} | {(p['rd'] + p['wr']) / p['n'] / 1e6:.2f} |\n")
print(open(dst).read()[:7000])
if len(sys.argv) > 3:
    fam_map = {"gemm_bf16": "_kernel<16-bit", "gemm_f16": "_kernel<16-bit", "absmax_scale": "ofq absmax_scale_kernel", "gemm_i8": "_kernel<i8",
               "qkr_attn_fwd": "ofq qkr_attn_fwd", "qkr_attn_bwd": "ofq qkr_attn_bwd_kernel", "gemm_lsq": "ofq gemm_lsq_kernel", "cga_adamw": "ofq cga_adamw", "adamw_multi": "ofq adamw_multi_kernel", "lsq_bwd": "ofq lsq_bwd_stream_kernel", "lsq_quant": "ofq lsq_quant_vec_kernel",
               "grad_prep": "ofq grad_prep_stream_kernel", "softmax_quant": "ofq softmax_quant_vec_kernel", "softmax_quant_bwd": "ofq softmax_quant_bwd_vec_kernel", "layernorm_bwd": "ofq layernorm_bwd_kernel", "layernorm_fwd": "ofq layernorm_fwd_kernel",
               "codes_to_bf16": "ofq codes_convert_kernel"}
    out = {}
    for name, pat in fam_map.items():
        sel = [p for k, p in per.items() if pat in k]
        n = sum(p["n"] for p in sel)
        if n:
            out[name] = {"dram_bytes_per_launch": sum(p["rd"] + p["wr"] for p in sel) / n, "launches": n, "source": src}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
