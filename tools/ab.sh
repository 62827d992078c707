#!/bin/bash
# A/B of two builds of the library inside ONE gpurun call (box-to-box variance is several percent):
#   tools/ab.sh [reps]   -> alternates libpqp_base.so (A) and libpqp_b200.so (B) on the cfg-2 perf case
reps=${1:-3}
for i in $(seq $reps); do
  for v in base b200; do
    PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_$v.so timeout 200 python tools/gpu_check.py perf 2>&1 | grep -o "t_solve_wall_s[^,]*\|solve_ms_dev[^,]*" | tr '\n' ' ' | sed "s/^/$v /"; echo
  done
done
