#!/bin/bash
# single GPU: per-sector L1 prefetch of the coming tile in k2a_v3 (K3_XPF=3) against the per-line prefetch
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/aa_bench_filters_base.log 2>&1
timeout 300 python bench.py $B > gpurun_out/aa_bench_base.log 2>&1
SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_xpf3.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/aa_bench_filters_xpf3.log 2>&1
SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_xpf3.so timeout 300 python bench.py $B > gpurun_out/aa_bench_xpf3.log 2>&1
SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_xpf3.so timeout 300 python bench.py --plan CBAND_143E --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/aa_bench_cband_xpf3.log 2>&1
timeout 300 python bench.py --plan CBAND_143E --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/aa_bench_cband_base.log 2>&1
echo done
