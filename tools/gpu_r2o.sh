#!/bin/bash
# Round 2, call O: end-to-end pipelines at B = 4096 (PQP_E2E=fused|chunks|plain) and at B = 1024.
set -u
mkdir -p gpurun_out
for B in 4096 1024; do for m in fused chunks plain fused chunks plain; do
  PQP_E2E=$m timeout 600 python bench.py --batch $B --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('B=$B mode=$m', 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
"
done; done 2>&1 | tee gpurun_out/e2e_modes.log
for m in fused plain; do echo "== breakdown $m"; PQP_E2E=$m python tools/e2e_breakdown.py 4096 2>&1 | tail -3 | tee -a gpurun_out/e2e_modes.log; done
