#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "16_callbacks or split_invariance or streams_are_independent or every_cta" > gpurun_out/k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/k_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
timeout 300 python bench.py $B > gpurun_out/k_bench.log 2>&1
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/k_bench_only_filters.log 2>&1
SDRB_K3_REGS=200 SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/k_bench_filters_r200.log 2>&1
SDRB_K3_REGS=200 timeout 300 python bench.py $B > gpurun_out/k_bench_r200.log 2>&1
tail -3 gpurun_out/k_tests.log
