#!/bin/bash
# call 3A: tsym_rank4 on the FP64 tensor cores (16 x 16 blocks of DMMA m8n8k4) A/B against the build before
# (libpqp_base.so), then the whole GPU suite on the new build and the reference arm with the sustained thread choice
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== cfg 2, B=4096 (base / new alternating)"; PERF_B=4096 bash tools/ab.sh 3 2>&1 | tee gpurun_out/ab_3a.log
echo "== GPU tests on the new build"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_3a.log
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/bench_ref_3a.json | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_ref_3a.json')); print({k:d[k] for k in ('value','steps_ms','threads_tried_ms','host_threads','cgroup_cpu_quota')}, d['cpu_baseline']['cores'])"
