#!/bin/bash
# Round 2, call B: A/B of the batched AXPY tail, full GPU test-suite, sweep of the large shapes on the new generic primitives.
set -u
mkdir -p gpurun_out
echo "== A/B (base = round-2 call A build, b200 = batched tail)"; PERF_B=4096 bash tools/ab.sh 2 2>&1 | tee gpurun_out/ab_tail.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== cfg sweep"; timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_b.log
