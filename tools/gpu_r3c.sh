#!/bin/bash
# call 3C: same-box comparison of builds (cfg 2, B = 4096): libpqp_base.so (rank-4 DMMA) against the libraries named in $LIBS
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/ab_3c.log
for i in 1 2 3; do
  for v in base $LIBS; do
    PERF_B=4096 PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_$v.so timeout 200 python tools/gpu_check.py perf 2>&1 | grep -o "solve_ms_dev[^,]*" | sed "s/^/$v /" | tee -a gpurun_out/ab_3c.log
  done
done
