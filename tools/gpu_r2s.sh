#!/bin/bash
# Round 2, call S: cost of the fallback's code on the big shapes (A/B against the build before it), full test-suite incl. slow set.
set -u
mkdir -p gpurun_out
for i in 1 2; do
  echo "== new"; SWEEP_FULL=1 timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_s.log
  echo "== base (before the whole-KKT fallback)"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_base.so SWEEP_FULL=1 timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_s.log
done
echo "== pytest, slow Maros-Meszaros problems included"; PQP_TEST_SLOW=1 timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_slow.log
