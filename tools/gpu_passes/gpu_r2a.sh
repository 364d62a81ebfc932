#!/bin/bash
# round 2, first GPU pass: parity tests, then A/B of the sub-VFO cascade kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/a_bench_v3.log 2>&1
SDRB_PER_CB=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/a_bench_v3_percb.log 2>&1
SDRB_K2A_V3=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/a_bench_v2.log 2>&1
SDRB_K2A_V3=0 SDRB_PER_CB=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/a_bench_v2_percb.log 2>&1
for w in 1480 4440 8880; do SDRB_K3_WARPS=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/a_bench_v3_w$w.log 2>&1; done
tail -3 gpurun_out/a_tests.log
# DC recursion alone / filters alone (profiling knobs; outputs meaningless)
SDRB_DEBUG_ONLY=dc timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/a_bench_only_dc.log 2>&1
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans > gpurun_out/a_bench_only_filters.log 2>&1
