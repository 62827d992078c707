#!/bin/bash
# call 3D: sweep inversion with the pivot phase one pass ahead (libpqp_b200.so) against the three-barrier form (libpqp_base.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== cfg 2, B=4096 (base / new alternating)"; PERF_B=4096 bash tools/ab.sh 3 2>&1 | tee gpurun_out/ab_3d.log
echo "== GPU tests on the new build"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_3d.log
echo "== stress (repeated launches)"; timeout 300 python tools/stress_launch.py 2>&1 | tail -3 | tee gpurun_out/stress_3d.log
