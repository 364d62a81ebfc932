#!/bin/bash
# single GPU: publish leg with 4 / 8 / 12 / 16 publisher sockets
mkdir -p gpurun_out
for n in 4 12 16 8; do
  SDRB_BENCH_SOCKETS=$n timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-plans > gpurun_out/z_bench_sock$n.log 2>&1
done
echo done
