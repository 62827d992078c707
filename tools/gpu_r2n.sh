#!/bin/bash
# Round 2, call N: does placing the pinned inputs on the GPU's NUMA node lift e2e? (same box, both ways), then gpu_round.sh
set -u
mkdir -p gpurun_out
for v in 1 0 1 0; do
  BENCH_NO_NUMA=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('no_numa=$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['config'].get('host_buffers'))
"
done 2>&1 | tee gpurun_out/ab_numa.log
python tools/e2e_breakdown.py 4096 2>&1 | tail -6 | tee -a gpurun_out/ab_numa.log
