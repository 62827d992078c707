#!/bin/bash
# Round 2, call M: A/B "base" = scalar block reductions de-inlined vs current; sweeps without the L2 set-aside.
set -u
mkdir -p gpurun_out
echo "== A/B cfg 2 (base = de-inlined block_max1 / block_sum1)"; PERF_B=4096 bash tools/ab.sh 3 2>&1 | tee gpurun_out/ab_m.log
echo "== cfg sweep, BASELINE batch sizes"; SWEEP_FULL=1 timeout 900 python tools/cfg_sweep.py 2b 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_m.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
