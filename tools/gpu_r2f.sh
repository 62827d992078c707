#!/bin/bash
# Round 2, call F: big variant with 8 loads in flight in the packed rank updates; full-size sweeps; new tests.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader > gpurun_out/clocks_f.txt
echo "== pytest (changed parts)"; timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_parity.py -m gpu -q -x -k "not maros and not repeated and not full_size" 2>&1 | tail -6 | tee gpurun_out/pytest_f.log
echo "== cfg sweep (profile)"; PQP_PROFILE=1 timeout 600 python tools/cfg_sweep.py 4 5 2>&1 | tee gpurun_out/cfg_sweep_f.log
echo "== cfg sweep, BASELINE batch sizes"; SWEEP_FULL=1 timeout 1200 python tools/cfg_sweep.py 2b 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_f.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader >> gpurun_out/clocks_f.txt
echo "== maros (rest of the list)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "maros" 2>&1 | tail -6 | tee gpurun_out/pytest_maros.log
