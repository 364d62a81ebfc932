#!/bin/bash
# 8-GPU box: concurrent host-link probe at 1/2/4/8 ranks, then the bench at 8 ranks (kernel-resident and end-to-end legs)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n_topo.log 2>&1
lscpu | head -25 >> gpurun_out/n_topo.log 2>&1
numactl -H >> gpurun_out/n_topo.log 2>&1
P=29511
for n in 1 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((P+n)) tools/pcie_probe_n.py > gpurun_out/n_probe_$n.json 2> gpurun_out/n_probe_$n.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+20)) bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-plans --no-zmq > gpurun_out/n_bench8.log 2>&1
tail -c 600 gpurun_out/n_probe_8.json
