#!/bin/bash
# Round 2, call C: A/B (wall time incl. retries) of tail batching + capacity 114 + skip of unchanged active sets;
# phase profile of the large shapes; GPU tests of the changed general kernel.
set -u
mkdir -p gpurun_out
echo "== A/B"; PERF_B=4096 bash tools/ab.sh 2 2>&1 | tee gpurun_out/ab_c.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== cfg sweep with phase profile"; PQP_PROFILE=1 timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_c.log
echo "== cfg sweep"; timeout 900 python tools/cfg_sweep.py 2b 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_c.log
