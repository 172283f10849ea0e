"""Key metrics + top stall lines of an `ncu --set full` report -> markdown. usage: ncu_summary.py rep.ncu-rep out.md"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
units = rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]
md = [f"# ncu --set full: {rep}\n", "| metric | value | unit |", "|---|---:|---|"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        md.append(f"| {w} | {vals[i]} | {units[i]} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ci, si = h.index("# Samples"), h.index("Source")
stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
data = []
agg = {}
for r in rows[2:]:
    try:
        v = int(r[ci])
    except Exception:
        continue
    data.append((v, r))
    for i in stalls:
        if r[i].isdigit():
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
tot = sum(v for v, _ in data) or 1
md += ["", f"Warp-stall samples: {tot}. By reason: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]), "",
       "| samples % | SASS | top stall |", "|---:|---|---|"]
for v, r in sorted(data, key=lambda t: -t[0])[:15]:
    st = max(((int(r[i]) if r[i].isdigit() else 0, h[i]) for i in stalls))
    md.append(f"| {100 * v / tot:.1f} | `{r[si][:90]}` | {st[1]} |")
open(out, "w").write("\n".join(md) + "\n")
print("\n".join(md[:24]))
