#!/bin/bash
# call V: fewer barriers per insertion / deletion (8 -> 5, 7 -> 3): A/B against the build before (libpqp_base.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== cfg 2, B=4096 (base / new alternating)"; PERF_B=4096 bash tools/ab.sh 3 2>&1 | tee gpurun_out/ab_v.log
for rep in 1 2; do
  echo "== new"; SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_v.log
  echo "== base"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_base.so SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_v.log
done
echo "== GPU tests on the new build"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_v.log
echo "== racecheck, tile + big"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | tail -4 | tee gpurun_out/racecheck_v.log
PQP_LAYOUT=big timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | tail -4 | tee -a gpurun_out/racecheck_v.log
