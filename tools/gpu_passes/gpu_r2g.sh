#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
timeout 300 python bench.py $B > gpurun_out/g_bench.log 2>&1
SDRB_DEBUG_ONLY=dc timeout 300 python bench.py $B > gpurun_out/g_bench_only_dc.log 2>&1
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/g_bench_only_filters.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench_full.log 2>&1
tail -3 gpurun_out/g_tests.log
