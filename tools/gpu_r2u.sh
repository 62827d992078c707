#!/bin/bash
# call U: big variant at THREE CTAs per SM (85 registers per thread) where the layout fits a third of an SM (cfg 3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
  echo "== 3 CTAs/SM (libpqp_occ3.so, PQP_BIG_CTAS=3)"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_occ3.so PQP_BIG_CTAS=3 SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 2>&1 | tee -a gpurun_out/cfg_sweep_u.log
  echo "== 2 CTAs/SM, 85-register build"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_occ3.so SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 2>&1 | tee -a gpurun_out/cfg_sweep_u.log
  echo "== 2 CTAs/SM, shipped build"; SWEEP_FULL=1 timeout 600 python tools/cfg_sweep.py 3 2>&1 | tee -a gpurun_out/cfg_sweep_u.log
done
echo "== occupancy seen by the runtime"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_occ3.so PQP_BIG_CTAS=3 timeout 300 python - <<'PY'
import sys; sys.path.insert(0, '.')
from proxsuite_b200 import proxqp as px
db = px.dense.DenseBatch(8, 100, 50, 50, box_constraints=True)
print(db.launch_config(), db.occupancy(False), db.occupancy(True))
PY
echo "== parity tests on the 3-CTA build"; PQP_B200_LIB=$PWD/proxsuite_b200/libpqp_occ3.so PQP_BIG_CTAS=3 timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -q -x -k "cfg3 or box or big" 2>&1 | tail -3
