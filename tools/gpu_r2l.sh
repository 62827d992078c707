#!/bin/bash
# Round 2, call L: block deletion / insertion (rank-4) + box-block skip: A/B on cfg 2, tests, sweeps.
set -u
mkdir -p gpurun_out
echo "== A/B cfg 2 (base: sequential bordering)"; PERF_B=4096 bash tools/ab.sh 2 2>&1 | tee gpurun_out/ab_l.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== cfg sweep, BASELINE batch sizes"; SWEEP_FULL=1 timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_l.log
echo "== base lib, same sweep"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_base.so SWEEP_FULL=1 timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_l.log
