#!/bin/bash
# Round 2, call A: baseline of the round-1 kernels on this round's tests + diagnostics.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
echo "== cpu arm diag"; timeout 400 python tools/cpu_arm_diag.py 2>&1 | tee gpurun_out/cpu_arm_diag.log | cut -c1-600
echo "== bench reference arm (first)"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>gpurun_out/bench_ref_err.log | tee gpurun_out/bench_ref.json | cut -c1-300
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== cfg sweep"; timeout 600 python tools/cfg_sweep.py 2b 3 4 5 2>&1 | tee gpurun_out/cfg_sweep.log
echo "== prof"; PQP_PROFILE=1 timeout 300 python tools/gpu_check.py prof 2>&1 | tail -3 | tee gpurun_out/prof.log
echo "== bench"; timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json | cut -c1-1500; tail -2 gpurun_out/bench_err.log
echo "== racecheck (tile kernel, 4 QPs n=20)"; timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | tail -5 | tee gpurun_out/racecheck.log
echo "== synccheck"; timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_target.py 2>&1 | tail -3 | tee gpurun_out/synccheck.log
