#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h_tests.log
for p in 54W_all 54W_288K; do
  timeout 300 python bench.py --plan $p --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/h_bench_$p.log 2>&1
  SDRB_FUSE_LATE=0 timeout 300 python bench.py --plan $p --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/h_bench_${p}_nofuse.log 2>&1
done
tail -3 gpurun_out/h_tests.log
