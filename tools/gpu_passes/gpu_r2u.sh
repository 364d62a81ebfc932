#!/bin/bash
# single GPU: staged input (bulk copy one tile ahead) for the 5-stage k2a_v3 again, now that the grids are wave-fitted
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/u_bench_filters_base.log 2>&1
timeout 300 python bench.py $B > gpurun_out/u_bench_base.log 2>&1
SDRB_K3_XS200=1 SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/u_bench_filters_xs200.log 2>&1
SDRB_K3_XS200=1 timeout 300 python bench.py $B > gpurun_out/u_bench_xs200.log 2>&1
SDRB_K3_REGS=232 SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/u_bench_filters_xs232.log 2>&1
SDRB_K3_REGS=232 timeout 300 python bench.py $B > gpurun_out/u_bench_xs232.log 2>&1
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "plan or cf32 or every" > gpurun_out/u_tests.log 2>&1
tail -n 3 gpurun_out/u_tests.log
echo done
