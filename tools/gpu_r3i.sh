#!/bin/bash
# call 3I (last GPU seconds of the round): B^T lam as an AXPY pass over the active rows (libpqp_ax.so) against the shipped
# build (libpqp_b200.so), two alternations, then the GPU suite on the candidate
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/ab_3i.log
for i in 1 2; do
  for v in b200 ax; do
    PERF_B=4096 PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_$v.so timeout 100 python tools/gpu_check.py perf 2>&1 | grep -o "solve_ms_dev[^,]*" | sed "s/^/$v /" | tee -a gpurun_out/ab_3i.log
  done
done
export PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_ax.so
echo "== GPU tests (ax)"; timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_3i.log
