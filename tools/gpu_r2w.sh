#!/bin/bash
# call W: big variant with padded packed rows and 16-byte pair loads / stores (two rows x four pairs in flight per lane)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
  echo "== new"; SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_w.log
  echo "== base"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_base.so SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_w.log
done
echo "== phase profile cfg 5 (new)"; PQP_PROFILE=1 timeout 600 python tools/cfg_sweep.py 5 2>&1 | tee -a gpurun_out/cfg_sweep_w.log
echo "== GPU tests on the new build"; PQP_TEST_SLOW=1 timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_w.log
echo "== memcheck + racecheck, big"; PQP_LAYOUT=big timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py 2>&1 | tail -3 | tee gpurun_out/sanitize_w.log
PQP_LAYOUT=big timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | tail -2 | tee -a gpurun_out/sanitize_w.log
