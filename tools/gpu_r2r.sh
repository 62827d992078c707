#!/bin/bash
# Round 2, call R: the whole-KKT inverse fallback of the big variant on the problems the dual-block path cannot solve.
set -u
mkdir -p gpurun_out
echo "== QSCORPIO / QFORPLAN / QSCAGR25 / QCAPRI (automatic switch)"; PQP_DEBUG_TRACE=0 PQP_WATCHDOG_MS=100000 timeout 600 python tools/mm_gpu_debug2.py QSCORPIO QFORPLAN QSCAGR25 QCAPRI 2>&1 | grep -v "^ \[\|^\[\[" | tee gpurun_out/mm_kkt.log
echo "== forced fallback on small shapes + baseline tests"; PQP_FORCE_KKT=1 timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k "big_variant" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -3
echo "== cfg sweep"; SWEEP_FULL=1 timeout 900 python tools/cfg_sweep.py 3 4 5 2>&1 | tee gpurun_out/cfg_sweep_r.log
