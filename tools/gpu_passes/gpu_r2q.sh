#!/bin/bash
# single GPU: parity tests, then A/B: wave-fitted k2a_v3 grids (SDRB_K3_WAVES), k1_v2 tiles per CTA, dc_blocks rewrite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_facade_gpu.py -m gpu -x -q > gpurun_out/q_tests.log 2>&1
tail -5 gpurun_out/q_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
for w in 1 2 3; do
  SDRB_K3_WAVES=$w SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/q_bench_filters_waves$w.log 2>&1
  SDRB_K3_WAVES=$w timeout 300 python bench.py $B > gpurun_out/q_bench_waves$w.log 2>&1
done
SDRB_K3_WARPS=2960 SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/q_bench_filters_old2960.log 2>&1
SDRB_K3_WARPS=2960 timeout 300 python bench.py $B > gpurun_out/q_bench_old2960.log 2>&1
for v in tpc5 tpc10; do
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_$v.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/q_bench_filters_$v.log 2>&1
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_$v.so timeout 300 python bench.py $B > gpurun_out/q_bench_$v.log 2>&1
done
SDRB_DEBUG_ONLY=dc timeout 300 python bench.py $B > gpurun_out/q_bench_dc.log 2>&1
echo done
