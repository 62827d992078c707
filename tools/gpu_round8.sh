#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu8.log
echo "== bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench8_err.log | tee gpurun_out/bench8_n2.json; tail -3 gpurun_out/bench8_err.log
