#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== stress"; timeout 200 python tools/stress_launch.py 30 1 2>&1 | tail -1 | tee gpurun_out/stress7.log
echo "== bench"; timeout 400 python bench.py 2>gpurun_out/bench7_err.log | tee gpurun_out/bench7.json; tail -2 gpurun_out/bench7_err.log
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-120
echo "== ncu fused kernel"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:pqp_solve_kernel_fused -c 1 -f -o gpurun_out/fused_r01c python tools/ncu_target.py 592 1 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
