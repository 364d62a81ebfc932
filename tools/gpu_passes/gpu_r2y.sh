#!/bin/bash
# two GPUs: the bench exactly as the driver launches it (default legs: e2e, publish leg, other plans with their canaries)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/y_bench2.log 2>&1
tail -c 1200 gpurun_out/y_bench2.log
echo done
