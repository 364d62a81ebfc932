#!/bin/bash
# single GPU: run-to-run spread of the default bench line (5 runs, kernel and end-to-end legs only)
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 300 python bench.py --no-cpu-baseline --no-plans --no-zmq > gpurun_out/am_bench_run$i.log 2>&1
done
echo done
