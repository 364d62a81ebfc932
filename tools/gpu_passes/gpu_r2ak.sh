#!/bin/bash
# single GPU: ncu --set full of the spectrum feed kernel (radix-8 FFT) as it runs in the CBAND_143E leg of the bench
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spectrum_feed|k_input_samples" -s 8 -c 6 -f -o gpurun_out/r02k_spectrum python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-zmq > gpurun_out/ak_ncu.log 2>&1
tail -n 2 gpurun_out/ak_ncu.log | cut -c1-200
echo done
