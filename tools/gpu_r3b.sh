#!/bin/bash
# call 3B: 8-pivot sweep + rank-8 DMMA update (libpqp_b200.so) against the 4-pivot sweep + rank-4 DMMA update (libpqp_base.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== cfg 2, B=4096 (base / new alternating)"; PERF_B=4096 bash tools/ab.sh 3 2>&1 | tee gpurun_out/ab_3b.log
echo "== GPU tests on the new build"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_3b.log
