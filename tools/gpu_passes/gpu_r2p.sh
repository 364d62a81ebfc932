#!/bin/bash
# single GPU: full GPU suite, then A/B of (a) the walk's bulk-copy prefetch, (b) the packed complex multiplies of k2a_v3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/p_tests.log 2>&1
tail -5 gpurun_out/p_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
for k in 0 1; do
  SDRB_DCW_BULK=$k SDRB_DEBUG_ONLY=dc timeout 300 python bench.py $B > gpurun_out/p_bench_dc_bulk$k.log 2>&1
  SDRB_DCW_BULK=$k timeout 300 python bench.py $B > gpurun_out/p_bench_bulk$k.log 2>&1
done
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/p_bench_filters.log 2>&1
SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_scalarcmul.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/p_bench_filters_scalarcmul.log 2>&1
SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_scalarcmul.so timeout 300 python bench.py $B > gpurun_out/p_bench_scalarcmul.log 2>&1
echo done
