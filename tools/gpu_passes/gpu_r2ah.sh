#!/bin/bash
# single GPU: callbacks per call 2 / 4 / 8 (how much of the step is launch granularity)
mkdir -p gpurun_out
for nb in 2 4 8; do
  SDRB_DEBUG_ONLY=filters timeout 300 python bench.py --blocks $nb --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/ah_bench_filters_nb$nb.log 2>&1
  timeout 300 python bench.py --blocks $nb --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/ah_bench_nb$nb.log 2>&1
done
echo done
