#!/bin/bash
# Round 2, call E: the BIG variant of the tile body on the GPU: parity tests, sweep of cfg 3 / 4 / 5 with the phase profile.
set -u
mkdir -p gpurun_out
echo "== pytest (baseline configs + big variant)"; timeout 1800 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_big.log
echo "== cfg sweep with phase profile"; PQP_PROFILE=1 timeout 1200 python tools/cfg_sweep.py 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_e.log
echo "== cfg sweep"; timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee -a gpurun_out/cfg_sweep_e.log
echo "== pytest (everything else)"; timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_baseline_configs.py 2>&1 | tail -8 | tee gpurun_out/pytest_rest.log
