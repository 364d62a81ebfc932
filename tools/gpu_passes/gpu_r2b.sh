#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
timeout 300 python bench.py $B > gpurun_out/b_bench_v3.log 2>&1
SDRB_K2A_V3=0 timeout 300 python bench.py $B > gpurun_out/b_bench_v2.log 2>&1
SDRB_DEBUG_ONLY=dc timeout 300 python bench.py $B > gpurun_out/b_bench_only_dc.log 2>&1
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/b_bench_only_filters.log 2>&1
for w in 1480 5920; do SDRB_K3_WARPS=$w timeout 300 python bench.py $B > gpurun_out/b_bench_v3_w$w.log 2>&1; done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench_full.log 2>&1
tail -3 gpurun_out/b_tests.log
