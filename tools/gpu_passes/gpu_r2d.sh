#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "16_callbacks or dc_state or split_invariance or overlapped" > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/d_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
timeout 300 python bench.py $B > gpurun_out/d_bench.log 2>&1
SDRB_DCW_RING=4 timeout 300 python bench.py $B > gpurun_out/d_bench_ring4.log 2>&1
SDRB_DEBUG_ONLY=dc timeout 300 python bench.py $B > gpurun_out/d_bench_only_dc.log 2>&1
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/d_bench_only_filters.log 2>&1
SDRB_K2A_V3=0 timeout 300 python bench.py $B > gpurun_out/d_bench_v2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2a_v3|k0_dc_walk" -s 36 -c 6 -f -o gpurun_out/r02b_full python bench.py --steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/d_ncu.log 2>&1
tail -3 gpurun_out/d_tests.log
