#!/bin/bash
mkdir -p gpurun_out
SDRB_K3_WS=1 timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "16_callbacks or split_invariance or streams_are_independent" > gpurun_out/l_tests_ws.log 2>&1; echo "tests rc=$?" >> gpurun_out/l_tests_ws.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
for w in 1 2; do
  SDRB_K3_WS=$w SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/l_bench_filters_ws$w.log 2>&1
  SDRB_K3_WS=$w timeout 300 python bench.py $B > gpurun_out/l_bench_ws$w.log 2>&1
done
SDRB_K3_REGS=200 SDRB_K3_XS200=1 SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/l_bench_filters_r200xs.log 2>&1
SDRB_K3_REGS=200 SDRB_K3_XS200=1 timeout 300 python bench.py $B > gpurun_out/l_bench_r200xs.log 2>&1
SDRB_K3_WS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2a_v3" -s 12 -c 2 -f -o gpurun_out/r02f_ws python bench.py --steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/l_ncu.log 2>&1
tail -3 gpurun_out/l_tests_ws.log
