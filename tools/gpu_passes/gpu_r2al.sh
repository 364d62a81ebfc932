#!/bin/bash
# single GPU: the full GPU suite on the final tree
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/al_tests.log 2>&1
tail -n 5 gpurun_out/al_tests.log
echo done
