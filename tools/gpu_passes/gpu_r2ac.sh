#!/bin/bash
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
SDRB_K3_REGS=168 SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/ac_bench_filters_r168.log 2>&1
SDRB_K3_REGS=168 timeout 300 python bench.py $B > gpurun_out/ac_bench_r168.log 2>&1
SDRB_K3_REGS=168 SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_unroll2.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/ac_bench_filters_r168_unroll2.log 2>&1
SDRB_K3_REGS=168 SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_unroll2.so timeout 300 python bench.py $B > gpurun_out/ac_bench_r168_unroll2.log 2>&1
echo done
