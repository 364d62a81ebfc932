#!/bin/bash
# single GPU: compute-sanitizer on the final tree
mkdir -p gpurun_out
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/aj_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/aj_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py quick > gpurun_out/aj_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/aj_racecheck.log
timeout 700 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_small.py quick > gpurun_out/aj_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/aj_synccheck.log
tail -n 3 gpurun_out/aj_memcheck.log gpurun_out/aj_racecheck.log gpurun_out/aj_synccheck.log
echo done
