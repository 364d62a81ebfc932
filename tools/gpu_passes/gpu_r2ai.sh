#!/bin/bash
# 8-GPU box, end of the round: the bench exactly as the driver launches it (default legs)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/ai_bench8.log 2>&1
tail -c 600 gpurun_out/ai_bench8.log
echo done
