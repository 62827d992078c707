#!/bin/bash
# Round 2, call K (2 GPUs): bench under torchrun (NCCL gather from device buffers), sharded C-ABI test on two real devices.
set -u
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus_k.txt
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 2>gpurun_out/bench2_err.log | tee gpurun_out/bench_n2.json | cut -c1-700; tail -3 gpurun_out/bench2_err.log
echo "== bench reference arm N=2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/bench_ref_n2.json | cut -c1-200
echo "== sharded tests on 2 devices"; timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k "sharded" 2>&1 | tail -3
echo "== cfg 4 (BASELINE: 8192 x n=256 over 8 GPUs = 1024 per GPU) through ShardedBatch on 2 GPUs"; timeout 600 python tools/sharded_cfg4.py 2>&1 | tail -3 | tee gpurun_out/sharded_cfg4.log
