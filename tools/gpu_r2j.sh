#!/bin/bash
# Round 2, call J: A/B of the L2 access-policy window on the workspace (same library, PQP_L2_PERSIST=0|1).
set -u
mkdir -p gpurun_out
for i in 1 2; do for P in 0 1; do
  PQP_L2_PERSIST=$P PERF_B=4096 timeout 200 python tools/gpu_check.py perf 2>&1 | grep -o "t_solve_wall_s[^,]*\|solve_ms_dev[^,]*" | tr '\n' ' ' | sed "s/^/persist=$P /"; echo
done; done 2>&1 | tee gpurun_out/ab_l2.log
for P in 0 1; do
  echo "== dram bytes, persist=$P"
  PQP_L2_PERSIST=$P PQP_E2E=plain timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:pqp_solve_kernel -c 1 python tools/ncu_target.py 4096 1 2>&1 | grep -E "dram__|lts__|gpu__time" | tee -a gpurun_out/ab_l2.log
done
echo "== cfg sweep, BASELINE batch sizes"; SWEEP_FULL=1 timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_j.log
