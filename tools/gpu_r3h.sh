#!/bin/bash
# call 3H: __syncwarp before the mirror stores (libpqp_b200.so) and the fast path for blocks of off-diagonal tiles
# (libpqp_fp.so) against the build of call 3Z (libpqp_base.so); GPU tests + racecheck on the fast-path build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/ab_3h.log
for i in 1 2 3; do
  for v in base b200 fp; do
    PERF_B=4096 PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_$v.so timeout 200 python tools/gpu_check.py perf 2>&1 | grep -o "solve_ms_dev[^,]*" | sed "s/^/$v /" | tee -a gpurun_out/ab_3h.log
  done
done
export PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_fp.so
echo "== GPU tests (fp)"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_3h.log
echo "== racecheck (fp)"; timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | tail -4 | tee gpurun_out/racecheck_3h.log
