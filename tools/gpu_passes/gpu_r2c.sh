#!/bin/bash
# ncu --set full of the sub-VFO cascade (v3), the DC walk and the ingest kernel: one launch each, steady state
mkdir -p gpurun_out
B="--steps 1 --warmup 6 --no-cpu-baseline --no-e2e --no-plans --no-zmq"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k2a_v3|k0_dc_walk|k1_v2|k2b_v2" -s 60 -c 6 -f -o gpurun_out/r02a_full python bench.py $B > gpurun_out/c_ncu.log 2>&1
tail -5 gpurun_out/c_ncu.log
