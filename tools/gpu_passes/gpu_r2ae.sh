#!/bin/bash
# single GPU: ncu --set full of the /5 FIR (54W_all) after the staging / compile-time tap changes
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_late_v2" -s 4 -c 1 -f -o gpurun_out/r02j_54w python bench.py --plan 54W_all --steps 1 --warmup 4 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/ae_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_late_v2" -s 4 -c 1 -f -o gpurun_out/r02j_288k python bench.py --plan 54W_288K --steps 1 --warmup 4 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/ae_ncu2.log 2>&1
echo done
