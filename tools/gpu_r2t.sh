#!/bin/bash
# call T: split storage of the big variant (leading rows of the packed S^-1 in the left-over shared memory): A/B by env switch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for rep in 1 2; do
  echo "== split"; SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_t.log
  echo "== PQP_NO_SPLIT=1"; PQP_NO_SPLIT=1 SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_t.log
done
echo "== phase profile, cfg 3"; PQP_PROFILE=1 timeout 300 python tools/cfg_sweep.py 3 2>&1 | tee -a gpurun_out/cfg_sweep_t.log
echo "== occupancy + baseline-config tests"; timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -q -x 2>&1 | tail -3
echo "== bench (no cpu baseline)"; timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_t_err.log | tee gpurun_out/bench_t.json | cut -c1-700
