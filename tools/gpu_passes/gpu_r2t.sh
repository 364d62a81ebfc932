#!/bin/bash
# single GPU, final state of the round: default bench, ncu --set full of one steady-state step, launch list of the bench command
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/t_bench_full.log 2>&1
tail -c 600 gpurun_out/t_bench_full.log
for m in filters dc; do
  SDRB_DEBUG_ONLY=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/t_bench_$m.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k0_dc|k1_v2|k2a_v3|k2b_v2|k3_carry" -s 102 -c 18 -f -o gpurun_out/r02h_full python bench.py --steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/t_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/t_launches.log 2>&1
echo done
