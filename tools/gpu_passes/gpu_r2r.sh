#!/bin/bash
# single GPU: A/B register cap 255 of the 5-stage k2a_v3 and CTA sizes of the short-cascade groups; compute-sanitizer on the round-2 kernels
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/r_bench_filters_base.log 2>&1
SDRB_K3_REGS=255 SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/r_bench_filters_regs255.log 2>&1
SDRB_K3_REGS=255 timeout 300 python bench.py $B > gpurun_out/r_bench_regs255.log 2>&1
for w in 3 6 12; do
  SDRB_K3_CTA_WARPS2=$w SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/r_bench_filters_cw2_$w.log 2>&1
  SDRB_K3_CTA_WARPS2=$w timeout 300 python bench.py $B > gpurun_out/r_bench_cw2_$w.log 2>&1
done
timeout 300 python bench.py $B > gpurun_out/r_bench_base.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py quick > gpurun_out/r_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r_racecheck.log
tail -3 gpurun_out/r_memcheck.log gpurun_out/r_racecheck.log
echo done
