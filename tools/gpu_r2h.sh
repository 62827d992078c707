#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== big"; PQP_DEBUG_TRACE=0 PQP_WATCHDOG_MS=8000 timeout 300 python tools/mm_gpu_debug2.py 2>&1 | tee gpurun_out/mm_debug2.log | tail -120
echo "== general"; PQP_NO_BIG=1 PQP_DEBUG_TRACE=0 PQP_WATCHDOG_MS=8000 timeout 300 python tools/mm_gpu_debug2.py QSCORPIO 2>&1 | tee -a gpurun_out/mm_debug2.log | tail -30
