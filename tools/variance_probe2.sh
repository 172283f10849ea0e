#!/bin/bash
# Which kernel families differ between the two run-to-run modes of the step?
for i in 1 2 3 4 5 6 7 8; do
  python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/var_$i.json
done
python - <<'PY'
import json
runs = []
for i in range(1, 9):
    d = json.loads(open(f"gpurun_out/var_{i}.json").read())
    runs.append((d["ms_per_step"], d["roofline"]["families_ms_per_step"], d["roofline"].get("instrumented_step_ms")))
runs.sort(key=lambda r: r[0])
print("ms:", [round(r[0], 3) for r in runs])
fast, slow = runs[0], runs[-1]
print("fast", fast[0], "slow", slow[0])
for k in sorted(fast[1], key=lambda k: -abs(slow[1].get(k, 0) - fast[1][k])):
    print(f"{k:24s} fast {fast[1][k]:8.4f} slow {slow[1].get(k, 0):8.4f} diff {slow[1].get(k, 0) - fast[1][k]:+.4f}")
PY
