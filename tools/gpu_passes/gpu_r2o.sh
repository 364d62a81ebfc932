#!/bin/bash
# single GPU: full GPU test suite on the current tree, then the bench with the DC walk's run solves on/off
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/o_tests.log 2>&1
tail -5 gpurun_out/o_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
for r in 1 2 4; do
  SDRB_DC_RUN=$r SDRB_DEBUG_ONLY=dc timeout 300 python bench.py $B > gpurun_out/o_bench_dc_run$r.log 2>&1
  SDRB_DC_RUN=$r timeout 300 python bench.py $B > gpurun_out/o_bench_run$r.log 2>&1
done
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/o_bench_filters.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/o_bench_full.log 2>&1
tail -c 1500 gpurun_out/o_bench_full.log
echo done
