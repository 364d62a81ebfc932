#!/bin/bash
# single GPU: full GPU suite after the /late FIR changes (fast staging path, compile-time taps per phase), then the 54W plans
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/x_tests.log 2>&1
tail -n 4 gpurun_out/x_tests.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
for p in 54W_all 54W_288K; do
  timeout 300 python bench.py --plan $p $B > gpurun_out/x_bench_$p.log 2>&1
  SDRB_LATE_GENERIC=1 timeout 300 python bench.py --plan $p $B > gpurun_out/x_bench_${p}_generic.log 2>&1
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/x_bench_25E.log 2>&1
echo done
