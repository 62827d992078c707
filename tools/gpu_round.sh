#!/bin/bash
# One gpurun call covering what a round needs from the GPU (~2.5 min of box time):
#   GPU parity tests, launch-pattern stress, bench (+ reference arm), ncu launch list, one ncu --set full capture.
# Everything lands in gpurun_out/ (keep it under 64 MiB: one .ncu-rep only).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
echo "== pytest"; PQP_TEST_SLOW=${PQP_TEST_SLOW:-0} timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== stress"; timeout 200 python tools/stress_launch.py 30 1 2>&1 | tail -1 | tee gpurun_out/stress.log
echo "== bench"; timeout 400 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json; tail -2 gpurun_out/bench_err.log
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/bench_ref.json
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-160
echo "== sanitizers (tile / general / big kernels, small shapes)"
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | tail -2 | tee gpurun_out/racecheck.log
PQP_LAYOUT=big timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | tail -2 | tee -a gpurun_out/racecheck.log
PQP_LAYOUT=big timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_target.py 2>&1 | tail -1 | tee gpurun_out/memcheck_big.log
timeout 400 compute-sanitizer --tool synccheck python tools/sanitize_target.py 2>&1 | tail -1 | tee gpurun_out/synccheck.log
echo "== BASELINE shapes (cfg 2b / 3 / 4 / 5 at their own batch sizes)"; SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 2b 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_final.log
if [ "${NCU_FULL:-1}" = "1" ]; then
  # (the cfg-4 report is exported to CSV on the box and dropped: gpurun_out/ is capped at 64 MiB)
  echo "== ncu --set full (big variant, cfg 4 shape, 296 QPs)"; PQP_E2E=plain timeout 900 ncu --set full --clock-control none -k regex:pqp_solve_kernel -c 1 -f -o gpurun_out/solve_big_cfg4 python tools/ncu_target.py 296 1 cfg4 2>&1 | tail -2
  ncu -i gpurun_out/solve_big_cfg4.ncu-rep --page raw --csv > gpurun_out/solve_big_cfg4_raw.csv 2>/dev/null; rm -f gpurun_out/solve_big_cfg4.ncu-rep
  echo "== ncu --set full (plain solve kernel, the bench launch)"; PQP_E2E=plain timeout 900 ncu --set full --clock-control none --import-source on -k regex:pqp_solve_kernel -c 1 -f -o gpurun_out/solve_full python tools/ncu_target.py 4096 1 2>&1 | tail -2
  ncu -i gpurun_out/solve_full.ncu-rep --page raw --csv > gpurun_out/solve_full_raw.csv 2>/dev/null
  # back in the container:  python tools/ncu_metrics_json.py gpurun_out/solve_full_raw.csv 4096 "<capture>"   (profiles/ncu_*.json, read by bench.py)
  #                         python tools/ncu_funcs.py gpurun_out/solve_full.ncu-rep proxsuite_b200/libpqp_b200.so --kernel tilek   (per-function attribution)
fi
