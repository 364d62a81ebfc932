#!/bin/bash
# single GPU: k2b_v2 with 5 / 6 resident CTAs per SM (register cap 102 / 85) against 4
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
for v in uvminb5 uvminb6; do
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_$v.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/ag_bench_filters_$v.log 2>&1
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_$v.so timeout 300 python bench.py $B > gpurun_out/ag_bench_$v.log 2>&1
done
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/ag_bench_filters_base.log 2>&1
timeout 300 python bench.py $B > gpurun_out/ag_bench_base.log 2>&1
echo done
