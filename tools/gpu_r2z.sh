#!/bin/bash
# call Z: the evidence files of the final build (bench both arms, launch list, two ncu captures; the cfg-4 capture is
# exported to CSV on the box and its report dropped: gpurun_out/ is capped at 64 MiB)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench"; timeout 400 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json | cut -c1-300
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-300
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; grep -c pqp_ gpurun_out/launches.csv
echo "== ncu --set full (big variant, cfg 4 shape, 296 QPs)"; PQP_E2E=plain timeout 900 ncu --set full --clock-control none -k regex:pqp_solve_kernel -c 1 -f -o gpurun_out/solve_big_cfg4 python tools/ncu_target.py 296 1 cfg4 2>&1 | tail -2
ncu -i gpurun_out/solve_big_cfg4.ncu-rep --page raw --csv > gpurun_out/solve_big_cfg4_raw.csv 2>/dev/null; ls -la gpurun_out/solve_big_cfg4.ncu-rep; rm -f gpurun_out/solve_big_cfg4.ncu-rep
echo "== ncu --set full (plain solve kernel, the bench launch)"; PQP_E2E=plain timeout 900 ncu --set full --clock-control none --import-source on -k regex:pqp_solve_kernel -c 1 -f -o gpurun_out/solve_full python tools/ncu_target.py 4096 1 2>&1 | tail -2
ncu -i gpurun_out/solve_full.ncu-rep --page raw --csv > gpurun_out/solve_full_raw.csv 2>/dev/null
du -sh gpurun_out
