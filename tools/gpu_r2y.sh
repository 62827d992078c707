#!/bin/bash
# call Y: software L2 prefetch distance of the big variant's streaming passes (env switch on one build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
  for pf in 0 1 2 4; do
    echo "== PQP_PREFETCH=$pf"; PQP_PREFETCH=$pf SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_y.log
  done
done
echo "== big-variant tests (default prefetch)"; timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -q -x 2>&1 | tail -3
