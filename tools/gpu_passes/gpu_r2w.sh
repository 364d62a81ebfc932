#!/bin/bash
# single GPU: ncu --set full of the /5 FIR with the fused mix (54W_all) and of k2b_v2 on CBAND_143E
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_late_v2|k2b_v2|k1_v2" -s 12 -c 6 -f -o gpurun_out/r02i_54w python bench.py --plan 54W_all --steps 1 --warmup 4 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/w_ncu.log 2>&1
tail -n 2 gpurun_out/w_ncu.log | cut -c1-300
echo done
