#!/bin/bash
# single GPU: /late FIR with immediate-offset staging: parity of the 54W plans, then their bench lines (4 and 5 CTAs per SM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "54W or golden or every_cta or silence or split" > gpurun_out/af_tests.log 2>&1
tail -n 3 gpurun_out/af_tests.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
for p in 54W_all 54W_288K; do
  timeout 300 python bench.py --plan $p $B > gpurun_out/af_bench_$p.log 2>&1
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_lvminb5.so timeout 300 python bench.py --plan $p $B > gpurun_out/af_bench_${p}_minb5.log 2>&1
done
echo done
