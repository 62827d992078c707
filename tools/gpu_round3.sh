#!/bin/bash
set -u
mkdir -p gpurun_out
for v in b200 gc base; do
  for i in 1 2; do
    PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_$v.so timeout 120 python tools/stress_launch.py 25 1 2>&1 | tail -2 | sed "s/^/[$v $i] /"
  done
done 2>&1 | tee gpurun_out/stress.log
echo "== memcheck of the launch pattern (new lib)"; timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/stress_launch.py 1 1 2>&1 | tail -8 | tee gpurun_out/memcheck_stress.log
echo "== bench"; timeout 300 python bench.py --steps 10 2>gpurun_out/bench3_err.log | tee gpurun_out/bench3.json; tail -3 gpurun_out/bench3_err.log
