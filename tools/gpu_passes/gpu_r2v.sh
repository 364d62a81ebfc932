#!/bin/bash
# single GPU: the restructured output copy of k2a_v3 (parity + time), and how the coming tile's input is brought in (K3_XPF builds)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_facade_gpu.py -m gpu -x -q > gpurun_out/v_tests.log 2>&1
tail -n 3 gpurun_out/v_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/v_bench_filters_base.log 2>&1
timeout 300 python bench.py $B > gpurun_out/v_bench_base.log 2>&1
for v in xpf1 xpf2; do
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_$v.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/v_bench_filters_$v.log 2>&1
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_$v.so timeout 300 python bench.py $B > gpurun_out/v_bench_$v.log 2>&1
done
SDRB_K3_REGS=255 SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_xpf2.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/v_bench_filters_xpf2_regs255.log 2>&1
SDRB_K3_REGS=255 SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_xpf2.so timeout 300 python bench.py $B > gpurun_out/v_bench_xpf2_regs255.log 2>&1
echo done
