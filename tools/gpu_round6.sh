#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== stress"; timeout 200 python tools/stress_launch.py 40 1 2>&1 | tail -2 | tee gpurun_out/stress6.log
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 400 python bench.py 2>gpurun_out/bench6_err.log | tee gpurun_out/bench6.json; tail -2 gpurun_out/bench6_err.log
echo "== bench ref arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench6_ref.json
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
