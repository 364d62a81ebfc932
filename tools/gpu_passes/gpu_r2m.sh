#!/bin/bash
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans"
for v in nopipe unroll1 unroll4; do
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_$v.so SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/m_bench_filters_$v.log 2>&1
  SDRB_LIB=$PWD/sdrreceiver_b200/variants/lib_$v.so timeout 300 python bench.py $B > gpurun_out/m_bench_$v.log 2>&1
done
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py $B > gpurun_out/m_bench_filters_base.log 2>&1
timeout 300 python bench.py $B > gpurun_out/m_bench_base.log 2>&1
# final-state ncu: every kernel of a step, and the launch list of the bench command
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k0_dc|k1_v2|k2a_v3|k2b_v2|k3_carry" -s 102 -c 18 -f -o gpurun_out/r02g_full python bench.py --steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/m_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/m_launches.log 2>&1
echo done
