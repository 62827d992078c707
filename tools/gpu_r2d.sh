#!/bin/bash
# Round 2, call D: full GPU test-suite, bench (+ reference arm), ncu launch list of the bench command, one --set full capture.
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== bench reference arm"; timeout 400 python bench.py --impl reference --steps 5 --warmup 3 2>gpurun_out/bench_ref_err.log | tee gpurun_out/bench_ref.json | cut -c1-400
echo "== bench"; timeout 900 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json | cut -c1-600; tail -2 gpurun_out/bench_err.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu --set full (plain solve kernel, 592 QPs)"; PQP_E2E=plain timeout 600 ncu --set full --clock-control none --import-source on -k regex:pqp_solve_kernel -c 1 -f -o gpurun_out/solve_r02 python tools/ncu_target.py 592 1 2>&1 | tail -2
