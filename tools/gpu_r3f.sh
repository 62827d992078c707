#!/bin/bash
# call 3F: block maxima through redux.sync (libpqp_b200.so) against the shuffle butterfly of nanmax (libpqp_base.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== cfg 2, B=4096 (base / new alternating)"; PERF_B=4096 bash tools/ab.sh 3 2>&1 | tee gpurun_out/ab_3f.log
echo "== new"; SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_3f.log
echo "== base"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_base.so SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_3f.log
echo "== GPU tests on the new build"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_3f.log
