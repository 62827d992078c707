#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== memcheck plain (standalone set-up kernel)"; PQP_E2E=plain timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/ncu_target.py 32 1 2>&1 | tail -12 | tee gpurun_out/memcheck_plain.log
echo "== memcheck fused (gated feed)"; timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/ncu_target.py 300 1 2>&1 | tail -12 | tee gpurun_out/memcheck_fused.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench plain"; PQP_E2E=plain timeout 300 python bench.py --no-cpu-baseline --steps 5 2>gpurun_out/bench_plain_err.log | tee gpurun_out/bench_plain.json; tail -5 gpurun_out/bench_plain_err.log
echo "== bench fused"; timeout 300 python bench.py --no-cpu-baseline --steps 10 2>gpurun_out/bench_fused_err.log | tee gpurun_out/bench_fused2.json
echo "== A/B <0> vs <1> on resident data"
for i in 1 2; do
  timeout 200 python tools/gpu_check.py perf 2>&1 | grep -o "qps_per_s_dev[^,]*" | sed "s/^/k0 /"
  PQP_FORCE_FUSED_KERNEL=1 timeout 200 python tools/gpu_check.py perf 2>&1 | grep -o "qps_per_s_dev[^,]*" | sed "s/^/k1 /"
done 2>&1 | tee gpurun_out/ab_fused.log
echo "== ncu set-up kernel"; PQP_E2E=plain timeout 400 ncu --set full --clock-control none --import-source on -k regex:pqp_setup_kernel -c 1 -f -o gpurun_out/setup_r01b python tools/ncu_target.py 1024 1 2>&1 | tail -3
echo "== ncu fused kernel"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:pqp_solve_kernel -c 1 -f -o gpurun_out/fused_r01b python tools/ncu_target.py 592 1 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
