#!/bin/bash
# call 3E: big variant, rank-4 update on the tensor cores (libpqp_b200.so) against the scalar pair form (libpqp_base.so),
# BASELINE shapes 3 / 4 / 5, two alternations; then the GPU tests of the BASELINE configs + Maros-Meszaros on the new build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/cfg_sweep_3e.log
for rep in 1 2; do
  echo "== new" | tee -a gpurun_out/cfg_sweep_3e.log; SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_3e.log
  echo "== base" | tee -a gpurun_out/cfg_sweep_3e.log; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_base.so SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_3e.log
done
echo "== GPU tests on the new build"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_3e.log
