#!/bin/bash
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_xpf4.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/ab_bench_filters_xpf4.log 2>&1
SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_xpf4.so timeout 300 python bench.py $B > gpurun_out/ab_bench_xpf4.log 2>&1
echo done
