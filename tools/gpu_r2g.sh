#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== maros debug"; timeout 900 python tools/mm_gpu_debug.py 2>&1 | tee gpurun_out/mm_debug.log | tail -40
echo "== cfg sweep (profile)"; PQP_PROFILE=1 timeout 600 python tools/cfg_sweep.py 4 5 2>&1 | tee gpurun_out/cfg_sweep_g.log
echo "== cfg sweep, BASELINE batch sizes"; SWEEP_FULL=1 timeout 1200 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_g.log
