#!/bin/bash
# Round 2, call P: after reserving the fused kernel's static shared memory in the layout budget (two CTAs per SM again).
set -u
mkdir -p gpurun_out
for B in 4096 1024; do for m in fused chunks fused chunks; do
  PQP_E2E=$m timeout 600 python bench.py --batch $B --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('B=$B mode=$m', 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
"
done; done 2>&1 | tee gpurun_out/e2e_modes_p.log
echo "== breakdown fused"; python tools/e2e_breakdown.py 4096 2>&1 | tail -3 | tee -a gpurun_out/e2e_modes_p.log
echo "== cfg 2b"; timeout 300 python tools/cfg_sweep.py 2b | tee -a gpurun_out/e2e_modes_p.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
