#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
timeout 300 python bench.py $B > gpurun_out/j_bench.log 2>&1
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/j_bench_only_filters.log 2>&1
for r in 168 200; do SDRB_K3_REGS=$r SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/j_bench_filters_r$r.log 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2a_v3" -s 12 -c 2 -f -o gpurun_out/r02e_full python bench.py --steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/j_ncu.log 2>&1
tail -3 gpurun_out/j_tests.log
