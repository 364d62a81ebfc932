#!/bin/bash
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
SDRB_K1_BULK=1 timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "16_callbacks or split_invariance or dc_state" > gpurun_out/i_tests_bulk.log 2>&1; echo "tests rc=$?" >> gpurun_out/i_tests_bulk.log
for v in 0 1; do
  SDRB_K1_BULK=$v SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/i_bench_filters_bulk$v.log 2>&1
  SDRB_K1_BULK=$v timeout 300 python bench.py $B > gpurun_out/i_bench_bulk$v.log 2>&1
done
for w in 1 3 4; do SDRB_K3_CTA_WARPS=$w SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/i_bench_filters_cw$w.log 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_v2" -s 24 -c 2 -f -o gpurun_out/r02d_k1 python bench.py --steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/i_ncu.log 2>&1
SDRB_K1_BULK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_v2" -s 24 -c 2 -f -o gpurun_out/r02d_k1_bulk python bench.py --steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq >> gpurun_out/i_ncu.log 2>&1
tail -3 gpurun_out/i_tests_bulk.log
