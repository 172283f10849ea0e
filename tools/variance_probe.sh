#!/bin/bash
# Run-to-run spread of the device-timed step (same box, separate processes), default allocator vs expandable segments.
for conf in "" "expandable_segments:True"; do
  for i in 1 2 3 4 5 6; do
    PYTORCH_CUDA_ALLOC_CONF=$conf python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('alloc=[$conf]', round(d['ms_per_step'], 3), 'ms', round(d['e2e']['value'], 1), 'img/s e2e')"
  done
done
