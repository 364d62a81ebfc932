#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "16_callbacks or split_invariance or streams_are_independent" > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
for r in 168 200 232; do SDRB_K3_REGS=$r SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/f_bench_filters_r$r.log 2>&1; done
for r in 168 200 232; do SDRB_K3_REGS=$r timeout 300 python bench.py $B > gpurun_out/f_bench_r$r.log 2>&1; done
tail -3 gpurun_out/f_tests.log
