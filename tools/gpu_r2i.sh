#!/bin/bash
# Round 2, call I: A/B of W-in-shared-memory, full GPU test-suite, bench both arms, ncu launch list + --set full on the bench-sized launch.
set -u
mkdir -p gpurun_out
echo "== A/B (base: W in the L2 workspace)"; PERF_B=4096 bash tools/ab.sh 2 2>&1 | tee gpurun_out/ab_i.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench reference arm"; timeout 400 python bench.py --impl reference --steps 5 --warmup 3 2>gpurun_out/bench_ref_err.log | tee gpurun_out/bench_ref.json | cut -c1-300
echo "== bench"; timeout 900 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json | cut -c1-400; tail -2 gpurun_out/bench_err.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu --set full (plain solve kernel, the bench-sized launch: 4096 QPs)"; PQP_E2E=plain timeout 900 ncu --set full --clock-control none --import-source on -k regex:pqp_solve_kernel -c 1 -f -o gpurun_out/solve_r02b python tools/ncu_target.py 4096 1 2>&1 | tail -2
