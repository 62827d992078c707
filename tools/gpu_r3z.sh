#!/bin/bash
# final call of round 2: evidence of the shipped build (GPU tests, bench both arms, launch list, one ncu --set full capture of
# the bench launch, BASELINE shapes, racecheck of the tile kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== GPU tests"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_3z.log
echo "== bench"; timeout 400 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json | cut -c1-300
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-300
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; grep -c pqp_ gpurun_out/launches.csv
echo "== BASELINE shapes"; SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 2b 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_3z.log
echo "== ncu --set full (plain solve kernel, the bench launch)"; PQP_E2E=plain timeout 900 ncu --set full --clock-control none --import-source on -k regex:pqp_solve_kernel -c 1 -f -o gpurun_out/solve_full python tools/ncu_target.py 4096 1 2>&1 | tail -2
ncu -i gpurun_out/solve_full.ncu-rep --page raw --csv > gpurun_out/solve_full_raw.csv 2>/dev/null
echo "== racecheck (tile kernel, general kernel with box, diagonal Hessian)"; timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_target.py 2>&1 | tail -4 | tee gpurun_out/racecheck_3z.log
du -sh gpurun_out
