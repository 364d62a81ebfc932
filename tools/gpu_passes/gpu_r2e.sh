#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/e_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
timeout 300 python bench.py $B > gpurun_out/e_bench.log 2>&1
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/e_bench_only_filters.log 2>&1
for w in 1480 5920; do SDRB_K3_WARPS=$w timeout 300 python bench.py $B > gpurun_out/e_bench_w$w.log 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2a_v3" -s 12 -c 2 -f -o gpurun_out/r02c_full python bench.py --steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/e_ncu.log 2>&1
tail -3 gpurun_out/e_tests.log
