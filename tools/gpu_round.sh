#!/bin/bash
# One gpurun call: GPU tests, bench (fused / chunks / plain end-to-end modes), A/B against libpqp_base.so,
# end-to-end breakdown, ncu launch list. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench fused" ; timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_fused.json
echo "== bench chunks" ; PQP_E2E=chunks timeout 300 python bench.py --no-cpu-baseline --steps 5 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_chunks.json
echo "== bench plain" ; PQP_E2E=plain timeout 300 python bench.py --no-cpu-baseline --steps 5 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_plain.json
echo "== e2e breakdown" ; timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -8 | tee gpurun_out/e2e_breakdown.log
if [ -f proxsuite_b200/libpqp_base.so ]; then echo "== A/B value" ; bash tools/ab.sh 2 2>&1 | tee gpurun_out/ab.log ; fi
echo "== phase profile" ; PQP_PROFILE=1 timeout 200 python tools/gpu_check.py prof 2>&1 | tail -3 | tee gpurun_out/phase_profile.log
echo "== ncu launch list" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1 ; tail -2 gpurun_out/ncu_bench.log
